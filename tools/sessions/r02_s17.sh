#!/bin/bash
# round-2 GPU session 17: config 5 -- stack entries carry their box's entry distance; pops skip entries beyond the current hit (A/B against IPT_POP_CULL=0)
set -x
O=gpurun_out/r02_s17; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -4 $O/pytest_gpu.log
for v in shipped nocull; do
  lib=inverse_path_tracer_b200/lib/ab/$v.so; [ $v = shipped ] && lib=inverse_path_tracer_b200/lib/libipt_b200.so
  echo "== $v" >> $O/ab.txt
  IPT_TRACE_DEBUG=1 IPT_B200_LIB=$PWD/$lib timeout 600 python tools/stress_config5.py --reps 2 --cache-dir /tmp/ipt_stress 2>&1 | grep "render \[\|ipt:\|incoherent\|traversal work\|image ==" >> $O/ab.txt
  IPT_B200_LIB=$PWD/$lib timeout 600 python tools/stress_config5.py --only-render 4 --reps 2 --cache-dir /tmp/ipt_stress 2>&1 | grep "render_grad" >> $O/ab.txt
done
cat $O/ab.txt

#!/bin/bash
# round-2 GPU session 15: config 5 -- traversal kernel at 16 blocks/SM (32 registers), shade kernel without spills (templated, 2/3/4 blocks/SM)
set -x
O=gpurun_out/r02_s15; mkdir -p $O
for v in shipped t_minb16 w_minb3 w_minb2 tw_16_3; do
  lib=inverse_path_tracer_b200/lib/ab/$v.so; [ $v = shipped ] && lib=inverse_path_tracer_b200/lib/libipt_b200.so
  echo "== $v" >> $O/ab.txt
  IPT_TRACE_DEBUG=1 IPT_B200_LIB=$PWD/$lib timeout 600 python tools/stress_config5.py --only-render 2 --reps 3 --cache-dir /tmp/ipt_stress 2>&1 | grep "render \[\|ipt:" >> $O/ab.txt
  IPT_B200_LIB=$PWD/$lib timeout 600 python tools/stress_config5.py --only-render 4 --reps 2 --cache-dir /tmp/ipt_stress 2>&1 | grep "render_grad" >> $O/ab.txt
done
cat $O/ab.txt
for v in shipped t_minb16; do
  lib=inverse_path_tracer_b200/lib/ab/$v.so; [ $v = shipped ] && lib=inverse_path_tracer_b200/lib/libipt_b200.so
  echo "== $v" >> $O/incoherent.txt
  IPT_B200_LIB=$PWD/$lib timeout 600 python tools/stress_config5.py --only-render 3 --cache-dir /tmp/ipt_stress 2>&1 | grep "incoherent\|traversal work" >> $O/incoherent.txt
done
cat $O/incoherent.txt
timeout 1500 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -4 $O/pytest_gpu.log

#!/bin/bash
# round-2 GPU session 41: block size / resident blocks of the wavefront shading kernels on the final build; GPU suite of the final build
set -x
O=gpurun_out/r02_s41; mkdir -p $O
for v in shipped wb128m6 wb128m5 wb128m4 wb256m2; do
  lib=inverse_path_tracer_b200/lib/ab/$v.so; [ $v = shipped ] && lib=inverse_path_tracer_b200/lib/libipt_b200.so
  echo "== $v" >> $O/wf.txt
  IPT_B200_LIB=$PWD/$lib timeout 300 python tools/stress_config5.py --only-render 2 --reps 3 --cache-dir /tmp/ipt_stress 2>&1 | grep "render \[" | sed 's/{[^}]*}//' >> $O/wf.txt
  IPT_B200_LIB=$PWD/$lib timeout 300 python tools/stress_config5.py --only-render 4 --reps 2 --cache-dir /tmp/ipt_stress 2>&1 | grep "render_grad" >> $O/wf.txt
done
cat $O/wf.txt
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -3 $O/pytest_gpu.log

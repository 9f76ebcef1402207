#!/bin/bash
# round-2 GPU session 30: config 5 -- speculative traversal (a lane postpones its first leaf and keeps descending), IPT_SPECULATE=1 against the shipped build
set -x
O=gpurun_out/r02_s30; mkdir -p $O
for v in shipped spec; do
  lib=inverse_path_tracer_b200/lib/ab/$v.so; [ $v = shipped ] && lib=inverse_path_tracer_b200/lib/libipt_b200.so
  echo "== $v" >> $O/ab.txt
  IPT_B200_LIB=$PWD/$lib timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -2 >> $O/ab.txt
  IPT_B200_LIB=$PWD/$lib timeout 600 python tools/stress_config5.py --reps 3 --cache-dir /tmp/ipt_stress 2>&1 | grep "render \[wave\|incoherent\|traversal work\|image ==" >> $O/ab.txt
done
cat $O/ab.txt

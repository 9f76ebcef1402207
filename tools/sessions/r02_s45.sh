#!/bin/bash
# round-2 GPU session 45: leaf size 1, 2, 3
set -x
O=gpurun_out/r02_s45; mkdir -p $O
for v in leaf1 leaf2 leaf3; do
  lib=inverse_path_tracer_b200/lib/ab/$v.so; [ $v = shipped ] && lib=inverse_path_tracer_b200/lib/libipt_b200.so
  echo "== $v" >> $O/leaf.txt
  IPT_B200_LIB=$PWD/$lib timeout 600 python tools/stress_config5.py --reps 3 --cache-dir /tmp/ipt_stress 2>&1 | grep "render \[wave\|incoherent\|traversal work\|nodes=" | sed 's/{[^}]*}//' >> $O/leaf.txt
  IPT_B200_LIB=$PWD/$lib timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "closest_hit or adversarial or fuzz or wavefront" 2>&1 | tail -1 >> $O/leaf.txt
done
cat $O/leaf.txt

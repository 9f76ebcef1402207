#!/bin/bash
# round-2 GPU session 29: (distances bounded at 1e4) adversarial rays against the traversal's box test (far origins, axis-parallel directions, origins on vertices)
# for all three node-record formats: shipped (quantised binary), IPT_BVH_QUANT=2 (4-wide), IPT_BVH_QUANT=0 (float)
set -x
O=gpurun_out/r02_s29; mkdir -p $O
for v in shipped q4 float; do
  lib=inverse_path_tracer_b200/lib/ab/$v.so; [ $v = shipped ] && lib=inverse_path_tracer_b200/lib/libipt_b200.so
  echo "== $v" >> $O/adversarial.txt
  IPT_B200_LIB=$PWD/$lib timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "adversarial or closest_hit or fuzz or large_triangle" 2>&1 | tail -8 >> $O/adversarial.txt
done
cat $O/adversarial.txt

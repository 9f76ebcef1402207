#!/bin/bash
# round-2 GPU session 36: longest-paths-first work order for small launches (csrc/lpt.cu): parity, shard-sized launches with and without it, inverse at 1/8 shards
set -x
O=gpurun_out/r02_s36; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -6 $O/pytest_gpu.log
for lpt in 1 0; do
  echo "== IPT_LPT=$lpt: 512x512x64 unbounded, one of eight row shards (forward)" >> $O/lpt.txt
  IPT_LPT=$lpt timeout 300 python tools/profile_target.py --size 512 --spp 64 --depth 0 --reps 5 --row-step 8 2>&1 | grep "rep [1234]" | sed 's/{[^}]*}//' >> $O/lpt.txt
  echo "== IPT_LPT=$lpt: scene 0 500x500x64 unbounded, one of eight row shards (adjoint)" >> $O/lpt.txt
  IPT_LPT=$lpt timeout 300 python tools/profile_target.py --scene scene0 --size 500 --spp 64 --depth 0 --reps 5 --row-step 8 --mode grad 2>&1 | grep "rep [1234]" | sed 's/{[^}]*}//' >> $O/lpt.txt
  echo "== IPT_LPT=$lpt: in-library inverse iteration, shards 8 and 1" >> $O/lpt.txt
  IPT_LPT=$lpt timeout 300 python tools/inverse_overhead.py --shards 8 --iters 100 --impl library 2>&1 | tail -3 >> $O/lpt.txt
  IPT_LPT=$lpt timeout 300 python tools/inverse_overhead.py --shards 1 --iters 50 --impl library 2>&1 | tail -2 >> $O/lpt.txt
done
cat $O/lpt.txt

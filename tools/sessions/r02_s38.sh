#!/bin/bash
# round-2 GPU session 38: where does a 1/8 row shard of the 512x512x64 spp frame lose its time? (ncu of the shard-size launch; bounded vs unbounded depth)
set -x
O=gpurun_out/r02_s38; mkdir -p $O
for d in 0 8 16 32; do
  echo "== depth $d, row step 8, IPT_LPT=0" >> $O/depth.txt
  IPT_LPT=0 timeout 300 python tools/profile_target.py --size 512 --spp 64 --depth $d --reps 5 --row-step 8 2>&1 | grep "rep [1234]" | sed 's/{[^}]*}//' >> $O/depth.txt
  echo "== depth $d, whole frame" >> $O/depth.txt
  IPT_LPT=0 timeout 300 python tools/profile_target.py --size 512 --spp 64 --depth $d --reps 4 2>&1 | grep "rep [123]" | sed 's/{[^}]*}//' >> $O/depth.txt
done
cat $O/depth.txt
IPT_LPT=0 timeout 600 ncu --set full --clock-control none -k regex:k_paths -s 2 -c 1 -o $O/prof_shard python tools/profile_target.py --size 512 --spp 64 --depth 0 --reps 3 --row-step 8 > $O/ncu.log 2>&1

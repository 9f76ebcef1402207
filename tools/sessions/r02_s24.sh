#!/bin/bash
# round-2 GPU session 24: sample-group size vs launch tail -- the bench workload whole (row step 1) and as one of eight interleaved row shards (row step 8)
set -x
O=gpurun_out/r02_s24; mkdir -p $O
for g in 64 32 16 8; do
  echo "== group $g, whole image 1024x1024x1024spp" >> $O/groups.txt
  timeout 300 python tools/profile_target.py --size 1024 --spp 1024 --reps 3 --group $g 2>&1 | grep "rep " >> $O/groups.txt
  echo "== group $g, one of eight row shards (row step 8)" >> $O/groups.txt
  timeout 300 python tools/profile_target.py --size 1024 --spp 1024 --reps 3 --group $g --row-step 8 2>&1 | grep "rep " >> $O/groups.txt
done
cat $O/groups.txt

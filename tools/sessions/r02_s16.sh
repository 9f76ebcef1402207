#!/bin/bash
# round-2 GPU session 16: config 5 -- how the traversal kernel responds to resident blocks per SM (12 = register limit at 40 registers)
set -x
O=gpurun_out/r02_s16; mkdir -p $O
for nb in 12 10 8 6 4; do
  echo "== IPT_TRACE_BLOCKS=$nb" >> $O/occupancy.txt
  IPT_TRACE_BLOCKS=$nb IPT_TRACE_DEBUG=1 timeout 600 python tools/stress_config5.py --only-render 2 --reps 3 --cache-dir /tmp/ipt_stress 2>&1 | grep "render \[\|ipt:" >> $O/occupancy.txt
done
cat $O/occupancy.txt

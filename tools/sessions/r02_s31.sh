#!/bin/bash
# round-2 GPU session 31: DRAM traffic of the bench-size k_paths<RENDER> launch on the final build (sample groups of 32 -> 402 MB of partial sums)
set -x
O=gpurun_out/r02_s31; mkdir -p $O
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_paths -s 3 -c 1 --csv --log-file $O/traffic_bench_kernel.csv python bench.py --steps 1 --warmup 3 --no-extras --no-cpu > $O/ncu.log 2>&1
cat $O/traffic_bench_kernel.csv | tail -5

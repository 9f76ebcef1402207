#!/bin/bash
# round-2 GPU session 42: bench line of the final build (inner quorum 14)
set -x
O=gpurun_out/r02_s42; mkdir -p $O
timeout 1500 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"
grep -h -o '"value": [0-9.]*, "unit": "Mrays/s", "n_gpus": [0-9]*' $O/bench_n1.json

#!/bin/bash
# round-2 GPU session 27: bench.py at N = 1 with the library's default sample-group rule (cap 32)
set -x
O=gpurun_out/r02_s27; mkdir -p $O
timeout 1500 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"
grep -h -o '"value": [0-9.]*, "unit": "Mrays/s", "n_gpus": [0-9]*' $O/bench_n1.json

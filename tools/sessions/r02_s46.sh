#!/bin/bash
# round-2 GPU session 46: GPU suite + bench line of the final build (leaf size 2)
set -x
O=gpurun_out/r02_s46; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log; tail -3 $O/pytest_gpu.log
timeout 1500 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"
grep -h -o '"value": [0-9.]*, "unit": "Mrays/s", "n_gpus": [0-9]*' $O/bench_n1.json

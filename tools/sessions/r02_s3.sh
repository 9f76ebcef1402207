#!/bin/bash
# round-2 GPU session 3: full GPU test suite (incl. the in-library inverse loop and its two-rank exchange), fixed per-iteration cost
set -x
O=gpurun_out/r02_s3; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -30 $O/pytest_gpu.log
for sh in 1 8; do for impl in python library; do timeout 300 python tools/inverse_overhead.py --shards $sh --iters 100 --impl $impl --profile 0; done; done > $O/inverse_overhead.txt 2>&1
cat $O/inverse_overhead.txt

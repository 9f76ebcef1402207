#!/bin/bash
# round-2 GPU session 9: wavefront adjoint for BVH scenes (parity + config-5 timing), full bench line
set -x
O=gpurun_out/r02_s9; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -12 $O/pytest_gpu.log
timeout 900 python bench.py --steps 3 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"; tail -5 $O/bench_n1.err; cat $O/bench_n1.json

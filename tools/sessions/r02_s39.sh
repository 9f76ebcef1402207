#!/bin/bash
# round-2 GPU session 39: final build (after the two-stream extra in bench.py) -- GPU suite, smoke(), bench line (N=1) and its launch list
set -x
O=gpurun_out/r02_s39; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -rs > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -5 $O/pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
timeout 1500 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench.csv python bench.py --steps 2 --warmup 3 > $O/ncu_bench.log 2>&1
grep -h -o '"value": [0-9.]*, "unit": "Mrays/s", "n_gpus": [0-9]*' $O/bench_n1.json

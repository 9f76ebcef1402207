#!/bin/bash
# round-2 GPU session 25 (8 GPUs): sample-group cap 32 -- GPU suite on one GPU, then the driver's bench command at N = 1, 2, 4, 8
set -x
O=gpurun_out/r02_s25; mkdir -p $O
CUDA_VISIBLE_DEVICES=0 timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -3 $O/pytest_gpu.log
CUDA_VISIBLE_DEVICES=0 timeout 900 python bench.py --gpus 1 --steps 5 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err; echo "n1 rc=$?"
for n in 2 4 8; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 5 --warmup 3 > $O/bench_n$n.json 2> $O/bench_n$n.err; echo "n$n rc=$?"
done
grep -h -o '"value": [0-9.]*, "unit": "Mrays/s", "n_gpus": [0-9]*' $O/bench_n*.json

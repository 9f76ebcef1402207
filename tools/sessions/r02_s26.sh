#!/bin/bash
# round-2 GPU session 26 (8 GPUs): bench.py with the library's default sample-group rule (cap 32) at N = 2, 4, 8
set -x
O=gpurun_out/r02_s26; mkdir -p $O
for n in 8 4 2; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n --steps 5 --warmup 3 > $O/bench_n$n.json 2> $O/bench_n$n.err; echo "n$n rc=$?"
done
grep -h -o '"value": [0-9.]*, "unit": "Mrays/s", "n_gpus": [0-9]*' $O/bench_n*.json

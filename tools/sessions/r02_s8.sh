#!/bin/bash
# round-2 GPU session 8 (2 GPUs): the in-library inverse loop with its NVLink mailbox exchange between two processes
# (CUDA IPC), and the whole bench line under torchrun
set -x
O=gpurun_out/r02_s8; mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err; echo "bench rc=$?"
tail -5 $O/bench_n2.err; cat $O/bench_n2.json

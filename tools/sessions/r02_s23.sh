#!/bin/bash
# round-2 GPU session 23 (8 GPUs): the driver's scaling command at N=8 -- forward rows interleaved over ranks, in-library inverse with the NVLink mailbox exchange
set -x
O=gpurun_out/r02_s23; mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 5 --warmup 3 > $O/bench_n8.json 2> $O/bench_n8.err; echo "bench rc=$?"
tail -c 1500 $O/bench_n8.json; tail -5 $O/bench_n8.err

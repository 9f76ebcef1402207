#!/bin/bash
# round-2 GPU session 5: chunked launches (tail overlap) A/B, full bench line
set -x
O=gpurun_out/r02_s5; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -15 $O/pytest_gpu.log
for ch in 1 2 4; do for sh in 1 8; do echo "IPT_CHUNKS=$ch"; IPT_CHUNKS=$ch timeout 300 python tools/inverse_overhead.py --shards $sh --iters 100 --impl library --profile 0; done; done > $O/inverse_chunks.txt 2>&1
grep -v "^+" $O/inverse_chunks.txt
for ch in 1 2 4; do for rs in 1 8; do echo "IPT_CHUNKS=$ch row_step=$rs"; IPT_CHUNKS=$ch timeout 300 python tools/profile_target.py --size 512 --spp 64 --depth 0 --reps 4 --row-step $rs | tail -2; done; done > $O/render_chunks.txt 2>&1
grep -v "^+" $O/render_chunks.txt
timeout 900 python bench.py --steps 3 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"; tail -5 $O/bench_n1.err; cat $O/bench_n1.json

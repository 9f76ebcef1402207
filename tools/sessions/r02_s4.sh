#!/bin/bash
# round-2 GPU session 4: camera-ray ring + guided claims + new group rule: parity, A/B against the previous build, inverse loop
set -x
O=gpurun_out/r02_s4; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -15 $O/pytest_gpu.log
timeout 900 python tools/ab_variants.py --workloads 512x64,1024x256 --reps 5 > $O/ab_render.txt 2>&1
cat $O/ab_render.txt
timeout 600 python tools/ab_variants.py --scene scene0 --workloads 500x64 --reps 5 > $O/ab_render_scene0.txt 2>&1
timeout 600 python tools/ab_variants.py --scene scene0 --mode grad --workloads 500x64 --reps 5 > $O/ab_grad_scene0.txt 2>&1
timeout 600 python tools/ab_variants.py --scene scene0 --mode graph --workloads 500x64 --reps 5 > $O/ab_graph_scene0.txt 2>&1
cat $O/ab_render_scene0.txt $O/ab_grad_scene0.txt $O/ab_graph_scene0.txt
for sh in 1 8; do timeout 300 python tools/inverse_overhead.py --shards $sh --iters 100 --impl library --profile 0; done > $O/inverse_overhead.txt 2>&1
cat $O/inverse_overhead.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_paths -s 1 -c 1 -o $O/prof_render_s4 -f python tools/profile_target.py --size 512 --spp 64 --reps 2 > $O/ncu_render.log 2>&1
timeout 300 python tools/stress_config5.py --reps 2 > $O/stress.txt 2>&1; tail -20 $O/stress.txt

#!/bin/bash
# round-2 GPU session 7: per-item state in shared memory -> 78 registers -> six resident blocks per SM (A/B against five)
set -x
O=gpurun_out/r02_s7; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -8 $O/pytest_gpu.log
timeout 900 python tools/ab_variants.py --only shipped,f_prev,g_fast5 --workloads 512x64,1024x256 --reps 5 > $O/ab_render.txt 2>&1
cat $O/ab_render.txt
timeout 600 python tools/ab_variants.py --only shipped,f_prev,g_fast5 --scene scene0 --workloads 500x64 --reps 5 > $O/ab_render_scene0.txt 2>&1
timeout 600 python tools/ab_variants.py --only shipped,f_prev,g_fast5 --scene scene0 --mode graph --workloads 500x64 --reps 5 > $O/ab_graph_scene0.txt 2>&1
cat $O/ab_render_scene0.txt $O/ab_graph_scene0.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_paths -s 1 -c 1 -o $O/prof_render_s7 -f python tools/profile_target.py --size 512 --spp 64 --reps 2 > $O/ncu_render.log 2>&1

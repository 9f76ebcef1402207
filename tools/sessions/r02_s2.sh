#!/bin/bash
# round-2 GPU session 2: record-stream intersection loop + explicit shared addressing; wavefront A/B on the brute-force Cornell box
set -x
O=gpurun_out/r02_s2; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -3 $O/pytest_gpu.log
timeout 900 python tools/ab_variants.py --workloads 512x64,1024x256 --reps 5 > $O/ab_render.txt 2>&1
cat $O/ab_render.txt
timeout 600 python tools/ab_variants.py --only shipped,a_r1 --scene scene0 --workloads 500x64 --reps 5 > $O/ab_render_scene0.txt 2>&1
timeout 600 python tools/ab_variants.py --only shipped,a_r1 --scene scene0 --mode grad --workloads 500x64 --reps 5 > $O/ab_grad_scene0.txt 2>&1
timeout 600 python tools/ab_variants.py --only shipped,a_r1 --scene scene0 --mode graph --workloads 500x64 --reps 5 > $O/ab_graph_scene0.txt 2>&1
cat $O/ab_render_scene0.txt $O/ab_grad_scene0.txt $O/ab_graph_scene0.txt
# wavefront (separate generate / extend / shade / accumulate kernels, state in HBM) vs persistent megakernel, both with the
# brute-force shared-memory intersection, Cornell box
for integ in 1 2; do for w in "512 64" "1024 256"; do set -- $w; echo "integrator=$integ size=$1 spp=$2"; timeout 300 python tools/profile_target.py --integrator $integ --size $1 --spp $2 --reps 3; done; done > $O/cornell_wavefront_ab.txt 2>&1
cat $O/cornell_wavefront_ab.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_paths -s 1 -c 1 -o $O/prof_render_s2 -f python tools/profile_target.py --size 512 --spp 64 --reps 2 > $O/ncu_render.log 2>&1

#!/bin/bash
# round-2 GPU session 1: parity of the trimmed k_paths, A/B of occupancy / fast-math variants, Cornell wavefront-vs-megakernel,
# one ncu --set full capture, sanitizer runs
set -x
O=gpurun_out/r02_s1; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.draw --format=csv > $O/gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -3 $O/pytest_gpu.log
timeout 900 python tools/ab_variants.py --workloads 512x64,1024x256 --reps 5 > $O/ab_render.txt 2>&1
cat $O/ab_render.txt
# wavefront vs megakernel on the Cornell box (traversal_mode 2 = BVH path; all 18 triangles land in the large-triangle set)
for integ in 1 2; do for w in "512 64" "1024 256"; do set -- $w; echo "integrator=$integ size=$1 spp=$2"; timeout 300 python tools/profile_target.py --traversal 2 --integrator $integ --size $1 --spp $2 --reps 3; done; done > $O/cornell_wavefront_ab.txt 2>&1
cat $O/cornell_wavefront_ab.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_paths -s 1 -c 1 -o $O/prof_render_s1 -f python tools/profile_target.py --size 512 --spp 64 --reps 2 > $O/ncu_render.log 2>&1
IPT_B200_LIB=$PWD/inverse_path_tracer_b200/lib/ab/b_mb6.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_paths -s 1 -c 1 -o $O/prof_render_s1_mb6 -f python tools/profile_target.py --size 512 --spp 64 --reps 2 > $O/ncu_render_mb6.log 2>&1
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python __graft_entry__.py --smoke > $O/sanitizer_memcheck_smoke.log 2>&1; echo "memcheck rc=$?" >> $O/sanitizer_memcheck_smoke.log
tail -5 $O/sanitizer_memcheck_smoke.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python __graft_entry__.py --smoke > $O/sanitizer_racecheck_smoke.log 2>&1; echo "racecheck rc=$?" >> $O/sanitizer_racecheck_smoke.log
tail -5 $O/sanitizer_racecheck_smoke.log

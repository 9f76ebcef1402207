#!/bin/bash
# round-2 GPU session 6: full GPU suite (bench-config bands, reference ipt_cuda.py unmodified when staged through
# $IPT_REFERENCE_IPT_CUDA, Phong adjoint, 32-bit light pick, GCN pipeline), adjoint pre-reduction A/B
set -x
O=gpurun_out/r02_s6; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -rs > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -25 $O/pytest_gpu.log
timeout 600 python tools/ab_variants.py --only shipped,e_sweepmatch --scene scene0 --mode grad --workloads 500x64 --reps 6 > $O/ab_grad_sweepmatch.txt 2>&1
cat $O/ab_grad_sweepmatch.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_paths -s 1 -c 1 -o $O/prof_render_s6 -f python tools/profile_target.py --size 512 --spp 64 --reps 2 > $O/ncu_render.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_paths -s 3 -c 1 -o $O/prof_grad_s6 -f python tools/profile_target.py --scene scene0 --size 500 --spp 64 --depth 0 --mode grad --reps 2 > $O/ncu_grad.log 2>&1
IPT_B200_LIB=$PWD/inverse_path_tracer_b200/lib/ab/e_sweepmatch.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_paths -s 3 -c 1 -o $O/prof_grad_s6_match -f python tools/profile_target.py --scene scene0 --size 500 --spp 64 --depth 0 --mode grad --reps 2 > $O/ncu_grad_match.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu > $O/ncu_bench.log 2>&1

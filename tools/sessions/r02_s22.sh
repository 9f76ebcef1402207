#!/bin/bash
# round-2 GPU session 22: full GPU suite, bench line (N=1) + its launch list, config-5 numbers of the shipped build,
# and the missing evidence for the 4-wide result: ncu --set full of k_trace_rays_bvh (second bounce) for the shipped build and for IPT_BVH_QUANT=2
set -x
O=gpurun_out/r02_s22; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -4 $O/pytest_gpu.log
timeout 1500 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"; tail -c 600 $O/bench_n1.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench.csv python bench.py --steps 2 --warmup 3 > $O/ncu_bench.log 2>&1
IPT_TRACE_DEBUG=1 timeout 600 python tools/stress_config5.py --reps 2 --cache-dir /tmp/ipt_stress > $O/stress.txt 2>&1
timeout 600 python tools/stress_config5.py --only-render 4 --reps 2 --cache-dir /tmp/ipt_stress 2>&1 | grep render_grad >> $O/stress.txt
grep "render \[\|ipt:\|incoherent\|traversal work\|image ==\|render_grad\|iptBuildBVH" $O/stress.txt
for v in shipped q4; do
  lib=inverse_path_tracer_b200/lib/ab/$v.so; [ $v = shipped ] && lib=inverse_path_tracer_b200/lib/libipt_b200.so
  IPT_B200_LIB=$PWD/$lib timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_trace_rays_bvh" -s 1 -c 1 -o $O/prof_trace_$v python tools/stress_config5.py --only-render 2 --reps 1 --cache-dir /tmp/ipt_stress > $O/ncu_trace_$v.log 2>&1
done
ls -la $O

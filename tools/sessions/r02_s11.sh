#!/bin/bash
# round-2 GPU session 11: quantised BVH records with the clamped reciprocal, A/B against the float-record build
set -x
O=gpurun_out/r02_s11; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -6 $O/pytest_gpu.log
for lib in shipped h_bvhfloat; do echo "== $lib"; if [ $lib = shipped ]; then unset IPT_B200_LIB; else export IPT_B200_LIB=$PWD/inverse_path_tracer_b200/lib/ab/$lib.so; fi; timeout 600 python tools/stress_config5.py --reps 2 2>&1 | grep -v "device copy"; done > $O/stress_quant_ab.txt 2>&1
unset IPT_B200_LIB
grep -E "==|render \[|traversal work|incoherent|image ==" $O/stress_quant_ab.txt
timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"; tail -3 $O/bench_n1.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02_s11/bench_n1.json') if l.startswith('{')][0])
print(json.dumps(d['extra']['config5_room_sphere_5p2M_triangles'], indent=1))
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_trace_rays_bvh -s 4 -c 1 -o $O/prof_trace_bvh_s11 -f python tools/stress_config5.py --only-render 2 --reps 1 > $O/ncu_trace.log 2>&1

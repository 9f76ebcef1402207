#!/bin/bash
# round-2 GPU session 10: quantised 32-byte BVH records (A/B against the 64-byte float records), wavefront adjoint with
# warp-aggregated gradient atomics, config-5 stress
set -x
O=gpurun_out/r02_s10; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -12 $O/pytest_gpu.log
for q in 1 0; do echo "IPT_BVH_QUANT=$q"; IPT_BVH_QUANT=$q timeout 600 python tools/stress_config5.py --reps 2 2>&1 | grep -v "device copy"; done > $O/stress_quant_ab.txt 2>&1
cat $O/stress_quant_ab.txt
timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"; tail -3 $O/bench_n1.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02_s10/bench_n1.json') if l.startswith('{')][0])
print(json.dumps(d['extra']['config5_room_sphere_5p2M_triangles'], indent=1))
PY

#!/bin/bash
# round-2 GPU session 13: wavefront adjoint with per-warp privatised hot-triangle gradients
set -x
O=gpurun_out/r02_s13; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -6 $O/pytest_gpu.log
timeout 600 python tools/stress_config5.py --only-render 4 --reps 2 --cache-dir /tmp/ipt_stress > $O/adjoint_plain.txt 2>&1; tail -3 $O/adjoint_plain.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_wf|k_trace|k_resolve|k_loss|k_reduce" --csv --log-file $O/launches_adjoint.csv python tools/stress_config5.py --only-render 4 --reps 1 --cache-dir /tmp/ipt_stress > $O/ncu_adjoint.log 2>&1
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/r02_s13/launches_adjoint.csv')) if len(r)>5]
hdr=rows[0]; ik=hdr.index('Kernel Name'); iv=hdr.index('Metric Value'); 
agg=collections.OrderedDict()
for r in rows[1:]:
    name=r[ik].split('(')[0]
    try: v=float(r[iv].replace(',',''))
    except: continue
    a=agg.setdefault(name,[0,0.0]); a[0]+=1; a[1]+=v
for k,(n,t) in agg.items(): print("%-60s n=%4d total %10.3f ms"%(k[:60],n,t/1e6))
PY

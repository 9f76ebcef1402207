#!/bin/bash
# round-2 GPU session 33: ... and into k_wf_generate for the camera rays; suite + config-5 numbers + launch list
set -x
O=gpurun_out/r02_s33; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -4 $O/pytest_gpu.log
timeout 600 python tools/stress_config5.py --reps 3 --cache-dir /tmp/ipt_stress 2>&1 | grep "render \[\|incoherent\|image ==" > $O/stress.txt
timeout 600 python tools/stress_config5.py --only-render 4 --reps 2 --cache-dir /tmp/ipt_stress 2>&1 | grep render_grad >> $O/stress.txt
cat $O/stress.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_wf|k_trace|k_resolve" --csv --log-file $O/launches_frame.csv python tools/stress_config5.py --only-render 2 --reps 1 --cache-dir /tmp/ipt_stress > $O/ncu_frame.log 2>&1
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/r02_s33/launches_frame.csv')) if len(r)>5]
hdr=rows[0]; ik=hdr.index('Kernel Name'); iv=hdr.index('Metric Value')
agg=collections.OrderedDict()
for r in rows[1:]:
    try: v=float(r[iv].replace(',',''))
    except: continue
    a=agg.setdefault(r[ik].split('(')[0][:40],[0,0.0]); a[0]+=1; a[1]+=v
for k,(n,t) in agg.items(): print("%-42s n=%3d %8.3f ms"%(k,n,t/1e6))
PY

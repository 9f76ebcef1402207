#!/bin/bash
# round-2 GPU session 37: longest-paths-first order, heavy items spread one per claim; kernel-only times from a launch list
set -x
O=gpurun_out/r02_s37; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "longest or forward_image or row_sharding" > $O/pytest_gpu.log 2>&1; tail -3 $O/pytest_gpu.log
for lpt in 1 0; do
  echo "== IPT_LPT=$lpt: 512x512x64 unbounded, one of eight row shards (forward)" >> $O/lpt.txt
  IPT_LPT=$lpt timeout 300 python tools/profile_target.py --size 512 --spp 64 --depth 0 --reps 5 --row-step 8 2>&1 | grep "rep [1234]" | sed 's/{[^}]*}//' >> $O/lpt.txt
  echo "== IPT_LPT=$lpt: scene 0 500x500x64 unbounded, one of eight row shards (adjoint)" >> $O/lpt.txt
  IPT_LPT=$lpt timeout 300 python tools/profile_target.py --scene scene0 --size 500 --spp 64 --depth 0 --reps 5 --row-step 8 --mode grad 2>&1 | grep "rep [1234]" | sed 's/{[^}]*}//' >> $O/lpt.txt
  IPT_LPT=$lpt timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_lpt$lpt.csv python tools/profile_target.py --size 512 --spp 64 --depth 0 --reps 3 --row-step 8 > /dev/null 2>&1
done
cat $O/lpt.txt
python - <<'PY'
import csv
for l in (1,0):
    rows=[r for r in csv.reader(open('gpurun_out/r02_s37/launches_lpt%d.csv'%l)) if len(r)>5]
    hdr=rows[0]; ik=hdr.index('Kernel Name'); iv=hdr.index('Metric Value')
    print('IPT_LPT=%d'%l, [(r[ik].split('(')[0][-24:], round(float(r[iv].replace(',',''))/1e3,1)) for r in rows[1:] if 'Fill' not in r[ik]][-6:])
PY

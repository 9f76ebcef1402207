#!/bin/bash
# round-2 GPU session 19: config 5 -- float nodes (explicit shared stack addressing, seed with the ray) vs 32-byte quantised nodes with a one-instruction decode
set -x
O=gpurun_out/r02_s19; mkdir -p $O
for v in shipped q2; do
  lib=inverse_path_tracer_b200/lib/ab/$v.so; [ $v = shipped ] && lib=inverse_path_tracer_b200/lib/libipt_b200.so
  echo "== $v" >> $O/ab.txt
  IPT_B200_LIB=$PWD/$lib timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -2 >> $O/ab.txt
  IPT_TRACE_DEBUG=1 IPT_B200_LIB=$PWD/$lib timeout 600 python tools/stress_config5.py --reps 2 --cache-dir /tmp/ipt_stress 2>&1 | grep "render \[\|ipt:\|incoherent\|traversal work\|image ==" >> $O/ab.txt
done
cat $O/ab.txt

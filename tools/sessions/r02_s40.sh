#!/bin/bash
# round-2 GPU session 40: phase quorums of k_trace_rays_bvh re-swept on the final build (quantised records, 10 blocks/SM, fused brute-force pass)
set -x
O=gpurun_out/r02_s40; mkdir -p $O
for v in shipped iq6 iq8 iq12 iq14 iq16 fq4 fq12 fq16 rc64 rc256; do
  lib=inverse_path_tracer_b200/lib/ab/$v.so; [ $v = shipped ] && lib=inverse_path_tracer_b200/lib/libipt_b200.so
  echo "== $v" >> $O/quorum.txt
  IPT_B200_LIB=$PWD/$lib timeout 300 python tools/stress_config5.py --only-render 2 --reps 3 --cache-dir /tmp/ipt_stress 2>&1 | grep "render \[" | sed 's/{[^}]*}//' >> $O/quorum.txt
done
cat $O/quorum.txt

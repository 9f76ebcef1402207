#!/bin/bash
# round-2 GPU session 35: ncu --set full of k_wf_shade on the final build (templated, 80 registers, brute-force pass fused), config 5, second bounce
set -x
O=gpurun_out/r02_s35; mkdir -p $O
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_wf_shade" -s 1 -c 1 -o $O/prof_wf_shade_final python tools/stress_config5.py --only-render 2 --reps 1 --cache-dir /tmp/ipt_stress > $O/ncu_shade.log 2>&1
ls -la $O

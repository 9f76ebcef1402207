#!/bin/bash
# round-2 GPU session 14: ncu --set full of the config-5 wavefront kernels (second bounce's traversal and shade)
set -x
O=gpurun_out/r02_s14; mkdir -p $O
timeout 600 python tools/stress_config5.py --only-render 2 --reps 2 --cache-dir /tmp/ipt_stress > $O/plain.txt 2>&1; tail -3 $O/plain.txt
# launch order per frame: generate, then per bounce [trace small, trace bvh, shade]; skip the first frame (27+ launches) and take bounce 2 of the next
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_trace_rays_bvh" -s 1 -c 1 -o $O/prof_trace_bvh python tools/stress_config5.py --only-render 2 --reps 1 --cache-dir /tmp/ipt_stress > $O/ncu_trace.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_wf_shade" -s 1 -c 1 -o $O/prof_wf_shade python tools/stress_config5.py --only-render 2 --reps 1 --cache-dir /tmp/ipt_stress > $O/ncu_shade.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_wf|k_trace|k_resolve" --csv --log-file $O/launches_frame.csv python tools/stress_config5.py --only-render 2 --reps 1 --cache-dir /tmp/ipt_stress > $O/ncu_frame.log 2>&1
ls -la $O

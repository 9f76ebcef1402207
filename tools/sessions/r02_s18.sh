#!/bin/bash
# round-2 GPU session 18: config 5 -- leaf triangle record + id requested together, seed with the ray, explicit shared stack addressing, shade at 3 blocks/SM
set -x
O=gpurun_out/r02_s18; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -4 $O/pytest_gpu.log
IPT_TRACE_DEBUG=1 timeout 600 python tools/stress_config5.py --reps 2 --cache-dir /tmp/ipt_stress 2>&1 | grep "render \[\|ipt:\|incoherent\|traversal work\|image ==" > $O/stress.txt
timeout 600 python tools/stress_config5.py --only-render 4 --reps 2 --cache-dir /tmp/ipt_stress 2>&1 | grep "render_grad" >> $O/stress.txt
cat $O/stress.txt

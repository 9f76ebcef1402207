#!/bin/bash
# round-2 GPU session 43: measurement-only builds of k_paths<RENDER> -- what the three remaining scalar-half levers would buy if the canon changed
# (Philox4x32-7, no re-normalisation of the sampled direction, MUFU sin/cos); these builds do NOT reproduce the oracle's images
set -x
O=gpurun_out/r02_s43; mkdir -p $O
timeout 1200 python tools/ab_variants.py --workloads 512x64,1024x256 --reps 4 > $O/ab.txt 2>&1
cat $O/ab.txt

#!/usr/bin/env python
"""A/B of k_paths builds on one GPU: every lib/ab/*.so (inverse_path_tracer_b200.build.build_variant) plus the
shipped library runs tools/profile_target.py in its own process; prints the best kernel time per workload.

    python tools/ab_variants.py [--workloads 512x64,1024x256] [--reps 5] [--only name,name]
"""
import argparse
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ap = argparse.ArgumentParser()
ap.add_argument("--workloads", default="512x64,1024x256")
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--only", default="")
ap.add_argument("--scene", default="cornell")
ap.add_argument("--mode", default="render")
a = ap.parse_args()
libs = [("shipped", os.path.join(ROOT, "inverse_path_tracer_b200", "lib", "libipt_b200.so"))]
for f in sorted(glob.glob(os.path.join(ROOT, "inverse_path_tracer_b200", "lib", "ab", "*.so"))):
    libs.append((os.path.basename(f)[:-3], f))
if a.only:
    keep = set(a.only.split(","))
    libs = [l for l in libs if l[0] in keep]
print("%-22s %s" % ("variant", "  ".join("%16s" % w for w in a.workloads.split(","))))
for name, path in libs:
    cells = []
    for w in a.workloads.split(","):
        size, spp = w.split("x")
        env = dict(os.environ, IPT_B200_LIB=path)
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "profile_target.py"), "--scene", a.scene, "--mode", a.mode, "--size", size, "--spp", spp,
                            "--reps", str(a.reps)], env=env, capture_output=True, text=True)
        ms = [float(m) for m in re.findall(r"kernel ([0-9.]+) ms", r.stdout)]
        mr = [float(m) for m in re.findall(r"([0-9.]+) Mrays/s", r.stdout)]
        cells.append("%7.3f ms %6.0f" % (min(ms), max(mr)) if ms else "FAILED: " + (r.stderr.strip().splitlines() or ["?"])[-1][:60])
    print("%-22s %s" % (name, "  ".join("%16s" % c for c in cells)), flush=True)

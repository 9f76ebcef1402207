#!/usr/bin/env python
"""Small fixed workload for ncu captures (one GPU): renders the Cornell box / scene 0 a few times.

    python tools/profile_target.py [--scene cornell|scene0] [--size 512] [--spp 64] [--depth 8] [--reps 3] [--mode render|graph|grad|trace]
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import inverse_path_tracer_b200 as ipt  # noqa: E402
from inverse_path_tracer_b200.ipt_cuda import parse_scene_file  # noqa: E402

ASSETS = os.path.join(ROOT, "tests", "golden", "assets")
ap = argparse.ArgumentParser()
ap.add_argument("--scene", default="cornell")
ap.add_argument("--size", type=int, default=512)
ap.add_argument("--spp", type=int, default=64)
ap.add_argument("--depth", type=int, default=8)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--mode", default="render")
ap.add_argument("--traversal", type=int, default=0)
ap.add_argument("--group", type=int, default=-1)
ap.add_argument("--row-step", type=int, default=1)
ap.add_argument("--integrator", type=int, default=0, help="BVH scenes: 0 auto, 1 megakernel, 2 wavefront")
a = ap.parse_args()
if a.scene == "cornell":
    objs = [{"pos": [0, 0, 4], "ori": [0, 0, 0], "scl": [2, 2, 2], "obj": "./CornellBox/CornellBox-Empty-CO.obj", "mtl": "./CornellBox/CornellBox-Empty-CO.mtl"}]
else:
    objs = parse_scene_file(os.path.join(ASSETS, "scenes", "0.txt"))
sc = ipt.Scene(objs, asset_root=ASSETS, traversal_mode=a.traversal)
if a.integrator:
    sc.set_integrator(a.integrator)
rng = np.random.RandomState(0)
target = rng.randint(0, 256, (a.size, a.size, 3)).astype(np.uint8)
for i in range(a.reps):
    p = ipt.default_params(width=a.size, height=a.size, spp=a.spp, max_depth=a.depth, seed=i, group_samples=a.group, row_begin=0, row_end=a.size, row_step=a.row_step)
    if a.mode == "render":
        img, cnt = sc.render(p)
    elif a.mode == "graph":
        raw, data, cnt = sc.create_graph(p, target)
    elif a.mode == "grad":
        loss, grad, img = sc.render_grad(p, target)
        cnt = sc.last_counters()
    else:
        n = a.size * a.size * a.spp
        d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
        rays = np.concatenate([rng.uniform(-1.5, 1.5, (n, 3)) + [0, 0, 4], d], 1).astype(np.float32)
        tri, t = sc.trace(rays)
        cnt = {"paths": n, "ext_rays": n, "shadow_rays": 0}
    ms = sc.kernel_times_ms(1)
    rays = cnt["ext_rays"] + cnt["shadow_rays"]
    print("rep %d: %s  kernel %.3f ms  %.1f Mrays/s" % (i, cnt, ms[0] if len(ms) else float("nan"), rays / (ms[0] * 1e-3) / 1e6 if len(ms) else 0))

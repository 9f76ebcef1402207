#!/usr/bin/env python
"""Fixed per-iteration cost of the inverse loop: one rank's share of an N-GPU iteration (rows r, r+N, ...) run on
one GPU, timed with CUDA events and wall clock, plus a cProfile of the host side.

    python tools/inverse_overhead.py --shards 8 --iters 50
"""
import argparse
import cProfile
import os
import pstats
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import inverse_path_tracer_b200 as ipt  # noqa: E402
from inverse_path_tracer_b200.ipt_cuda import parse_scene_file  # noqa: E402
from inverse_path_tracer_b200.torch_api import TorchScene, render_loss  # noqa: E402

ASSETS = os.path.join(ROOT, "tests", "golden", "assets")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shards", type=int, default=8)
    ap.add_argument("--iters", type=int, default=50)
    ap.add_argument("--profile", type=int, default=1)
    ap.add_argument("--impl", default="python", choices=["python", "library"], help="python: render_loss + torch Adam per iteration; library: iptInverseStep")
    a = ap.parse_args()
    from PIL import Image
    dev = torch.device("cuda", 0)
    sc = TorchScene(parse_scene_file(os.path.join(ASSETS, "scenes", "0.txt")), asset_root=ASSETS, device=0)
    target = torch.from_numpy(np.asarray(Image.open(os.path.join(ROOT, "tests", "golden", "golden_0_true.png")).convert("RGB")).copy()).to(dev)
    if a.impl == "library":
        # the whole iteration inside the library (csrc/inverse.cu): one C call enqueues `iters` iterations
        from inverse_path_tracer_b200.torch_api import inverse_optimizer
        p = ipt.default_params(width=500, height=500, spp=64, max_depth=0, row_begin=0, row_end=500, row_step=a.shards)
        lopt = inverse_optimizer(sc, p, target, lr=0.02)
        lopt.set_kd(np.full((sc.nT, 3), 0.5, np.float32))
        lopt.step(1000, 5)
        lopt.state()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        lopt.step(0, a.iters)
        t_issue = time.perf_counter() - t0
        loss, _, steps = lopt.state()
        t_wall = time.perf_counter() - t0
        print("shards %d [library]: %.3f ms/iter wall (%.3f ms/iter host issue time), loss %.6f after %d steps" % (
            a.shards, 1e3 * t_wall / a.iters, 1e3 * t_issue / a.iters, loss, steps))
        ms = sc.scene.kernel_times_ms(4)
        print("last path-kernel launches (ms, newest first):", ["%.3f" % x for x in ms])
        return
    kd = torch.full((sc.nT, 3), 0.5, device=dev, requires_grad=True)
    opt = torch.optim.Adam([kd], lr=0.02)

    def it(i):
        opt.zero_grad(set_to_none=True)
        p = ipt.default_params(width=500, height=500, spp=64, max_depth=0, seed=i, row_begin=0, row_end=500, row_step=a.shards)
        loss = render_loss(kd, sc, p, target, reduce=False)
        loss.backward()
        opt.step()
        with torch.no_grad():
            kd.clamp_(1e-3, 1.0)
        return loss

    for i in range(5):
        it(1000 + i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for i in range(a.iters):
        it(i)
    t_issue = time.perf_counter() - t0
    e1.record()
    torch.cuda.synchronize()
    t_wall = time.perf_counter() - t0
    print("shards %d: %.3f ms/iter device (events), %.3f ms/iter host issue time, %.3f ms/iter wall" % (
        a.shards, e0.elapsed_time(e1) / a.iters, 1e3 * t_issue / a.iters, 1e3 * t_wall / a.iters))
    ms = sc.scene.kernel_times_ms(4)
    print("last path-kernel launches (ms, newest first):", ["%.3f" % x for x in ms])
    if a.profile:
        pr = cProfile.Profile()
        pr.enable()
        for i in range(a.iters):
            it(i)
        pr.disable()
        torch.cuda.synchronize()
        st = pstats.Stats(pr)
        st.sort_stats("cumulative").print_stats(18)


if __name__ == "__main__":
    main()

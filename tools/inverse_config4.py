#!/usr/bin/env python
"""BASELINE config 4: batch inverse over scenes/{0..99}.txt (same geometry, cube albedo ~ U(0,1)).  Targets are rendered
by this library's forward pass (500 x 500, 100 spp, seed = scene id; the reference's imgs/ are not shipped), every albedo
starts at 0.5, 200 Adam iterations of 64 spp each.  Scene i runs on GPU i mod N (one process per GPU under torchrun, no
collective); this script runs its share and reports the recovered cube albedo against the scene file's.

    python tools/inverse_config4.py [--scenes 100] [--iters 200]
    torchrun --nproc-per-node N tools/inverse_config4.py
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import inverse_path_tracer_b200 as ipt  # noqa: E402
from inverse_path_tracer_b200 import parallel  # noqa: E402
from inverse_path_tracer_b200.ipt_cuda import parse_scene_file  # noqa: E402
from inverse_path_tracer_b200.torch_api import TorchScene, render_loss  # noqa: E402

ASSETS = os.path.join(ROOT, "tests", "golden", "assets")
SEEN = [20, 21, 22, 23]   # cube triangles the camera sees well (tools/inverse_config3.py)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scenes", type=int, default=100)
    ap.add_argument("--iters", type=int, default=200)
    ap.add_argument("--lr", type=float, default=0.02)
    a = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    mine = parallel.scenes_for_rank(a.scenes, rank, world)
    errs_seen, errs_wall, losses = [], [], []
    tgt8 = torch.empty((500, 500, 3), dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for sid in mine:
        sc = TorchScene(parse_scene_file(os.path.join(ASSETS, "scenes", "%d.txt" % sid)), asset_root=ASSETS, device=local)
        truth = sc.materials().reshape(-1, 3).clone()
        sc.render(ipt.default_params(width=500, height=500, spp=100, seed=sid), out_u8=tgt8)
        kd = torch.full((sc.nT, 3), 0.5, device=dev, requires_grad=True)
        opt = torch.optim.Adam([kd], lr=a.lr)
        for it in range(a.iters):
            opt.zero_grad(set_to_none=True)
            loss = render_loss(kd, sc, ipt.default_params(width=500, height=500, spp=64, max_depth=0, seed=1000 + it), tgt8, reduce=False)
            loss.backward()
            opt.step()
            with torch.no_grad():
                kd.clamp_(1e-3, 1.0)
        err = (kd.detach() - truth).abs()
        errs_seen.append(float(err[SEEN].mean()))
        errs_wall.append(float(err[:16].mean()))
        losses.append(float(loss.detach()))
        sc.set_materials(truth)
        sc.close()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    n = len(mine)
    print("rank %d/%d: %d scenes x %d iterations in %.1f s = %.1f iterations/s; final loss mean %.2e; |albedo error| visible cube faces: mean %.4f, "
          "max %.4f; walls / floor / ceiling: mean %.4f" % (rank, world, n, a.iters, dt, n * a.iters / dt, float(np.mean(losses)), float(np.mean(errs_seen)),
                                                          float(np.max(errs_seen)), float(np.mean(errs_wall))))


if __name__ == "__main__":
    main()

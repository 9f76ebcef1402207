#!/usr/bin/env python
"""BASELINE config 3: recover the 30 x 3 albedos of scenes/0.txt from the reference's own render preds/0_true.png
(tests/golden/golden_0_true.png) -- every albedo starts at 0.5, Adam on the path adjoint, 64 spp per iteration,
seed = iteration index.  Prints the loss curve and the recovered albedos against the true ones (getMaterials).

    python tools/inverse_config3.py [--iters 200] [--lr 0.02] [--spp 64]
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
from inverse_path_tracer_b200.ipt_cuda import parse_scene_file  # noqa: E402
from inverse_path_tracer_b200.torch_api import TorchScene, render_loss  # noqa: E402

ASSETS = os.path.join(ROOT, "tests", "golden", "assets")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=200)
    ap.add_argument("--lr", type=float, default=0.02)
    ap.add_argument("--spp", type=int, default=64)
    a = ap.parse_args()
    from PIL import Image
    dev = torch.device("cuda", 0)
    sc = TorchScene(parse_scene_file(os.path.join(ASSETS, "scenes", "0.txt")), asset_root=ASSETS, device=0)
    truth = sc.materials().reshape(-1, 3).clone()
    target = torch.from_numpy(np.asarray(Image.open(os.path.join(ROOT, "tests", "golden", "golden_0_true.png")).convert("RGB")).copy()).to(dev)
    kd = torch.full((sc.nT, 3), 0.5, device=dev, requires_grad=True)
    opt = torch.optim.Adam([kd], lr=a.lr)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for it in range(a.iters):
        opt.zero_grad(set_to_none=True)
        loss = render_loss(kd, sc, ipt.default_params(width=500, height=500, spp=a.spp, max_depth=0, seed=it), target, reduce=False)
        loss.backward()
        opt.step()
        with torch.no_grad():
            kd.clamp_(1e-3, 1.0)
        if it % 25 == 0 or it == a.iters - 1:
            print("iter %3d  loss %.6f" % (it, float(loss.detach())))
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("%d iterations in %.2f s = %.1f iterations/s" % (a.iters, dt, a.iters / dt))
    got = kd.detach()
    err = (got - truth).abs()
    # triangles the camera or the light never sees carry no gradient and stay at their start value: report per surface
    names = ["floor"] * 2 + ["ceiling"] * 8 + ["back wall"] * 2 + ["right wall"] * 2 + ["left wall"] * 2 + ["light"] * 2 + ["cube"] * 12
    for surf in ("floor", "ceiling", "back wall", "right wall", "left wall", "light", "cube"):
        idx = [i for i, n in enumerate(names) if n == surf]
        print("%-10s true %s  recovered %s  mean |err| %.3f" % (surf, np.round(truth[idx].mean(0).cpu().numpy(), 3), np.round(got[idx].mean(0).cpu().numpy(), 3),
                                                             float(err[idx].mean())))
    cube = [i for i, n in enumerate(names) if n == "cube"]
    print("cube, per triangle |err| (faces the camera and the light never reach keep their start value 0.5):")
    print("  " + "  ".join("%d:%.3f" % (i, float(err[i].mean())) for i in cube))
    seen = [i for i in cube if float(err[i].mean()) < 0.05]
    if seen:
        print("  cube triangles recovered to < 0.05: %s -> recovered %s (true %s)" % (seen, np.round(got[seen].mean(0).cpu().numpy(), 3), np.round(truth[seen].mean(0).cpu().numpy(), 3)))
    print("mean |albedo error| over all 90 parameters: %.4f (start: %.4f)" % (float(err.mean()), float((0.5 - truth).abs().mean())))
    sc.close()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""BASELINE config 5 (stress): room = shapes/scene.obj with an emissive ceiling patch, plus shapes/sphere.obj
subdivided `--levels` more times with re-projection to r = 0.5 (6 levels -> 5 242 880 triangles), placed at
POS 0 -0.5 3, SCL 2.  Reports GPU BVH build time and closest-hit throughput on incoherent diffuse bounces.

    python tools/stress_config5.py --levels 6 --size 1024 --spp 16 --depth 8
"""
import argparse
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import inverse_path_tracer_b200 as ipt  # noqa: E402
from inverse_path_tracer_b200 import native  # noqa: E402

ASSETS = os.path.join(ROOT, "tests", "golden", "assets")


def read_obj(path):
    v, f = [], []
    for line in open(path):
        t = line.split()
        if not t:
            continue
        if t[0] == "v":
            v.append([float(x) for x in t[1:4]])
        elif t[0] == "f":
            f.append([int(x.split("/")[0]) - 1 for x in t[1:4]])
    return np.array(v, np.float64), np.array(f, np.int64)


def subdivide(v, f, r):
    """one 1-to-4 midpoint subdivision with re-projection to the sphere of radius r"""
    e = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]])
    e.sort(1)
    key = e[:, 0] * (v.shape[0] + 1) + e[:, 1]
    uk, inv = np.unique(key, return_inverse=True)
    a, b = uk // (v.shape[0] + 1), uk % (v.shape[0] + 1)
    mid = v[a] + v[b]
    mid *= r / np.linalg.norm(mid, axis=1, keepdims=True)
    nv = v.shape[0]
    v2 = np.concatenate([v, mid])
    m = inv.reshape(3, -1).T + nv   # per face: midpoints of edges (01, 12, 20)
    f2 = np.concatenate([np.stack([f[:, 0], m[:, 0], m[:, 2]], 1), np.stack([f[:, 1], m[:, 1], m[:, 0]], 1),
                         np.stack([f[:, 2], m[:, 2], m[:, 1]], 1), np.stack([m[:, 0], m[:, 1], m[:, 2]], 1)])
    return v2, f2


def write_obj(path, v, f, normals=True):
    with open(path, "w") as fh:
        fh.write("\n".join("v %.8f %.8f %.8f" % tuple(x) for x in v))
        fh.write("\n")
        if normals:
            n = v / np.linalg.norm(v, axis=1, keepdims=True)
            fh.write("\n".join("vn %.8f %.8f %.8f" % tuple(x) for x in n))
            fh.write("\n")
        fh.write("\n".join("f %d %d %d" % tuple(x) for x in (f + 1)))
        fh.write("\n")


def stress_scene_objects(levels, cache_dir=""):
    """write the config-5 assets (room with an emissive ceiling patch + subdivided sphere) and return
    (object list for ipt.Scene, sphere triangle count)"""
    tmp = cache_dir or tempfile.mkdtemp(prefix="ipt_stress_")
    os.makedirs(tmp, exist_ok=True)
    sphere = os.path.join(tmp, "sphere_big_%d.obj" % levels)
    if not os.path.exists(sphere):
        v, f = read_obj(os.path.join(ASSETS, "shapes", "sphere.obj"))
        for _ in range(levels):
            v, f = subdivide(v, f, 0.5)
        write_obj(sphere + ".tmp", v, f)
        os.replace(sphere + ".tmp", sphere)
    n_tris = 1280 * 4 ** levels
    # room: shapes/scene.obj; faces 17-18 (the ceiling patch) emissive through a synthetic MTL
    room_src = open(os.path.join(ASSETS, "shapes", "scene.obj")).read().splitlines()
    out, nf = ["mtllib room.mtl", "usemtl wall"], 0
    for line in room_src:
        if line.startswith("f "):
            nf += 1
            if nf == 17:
                out.append("usemtl lamp")
        out.append(line)
    room = os.path.join(tmp, "room.obj")
    open(room, "w").write("\n".join(out) + "\n")
    open(os.path.join(tmp, "room.mtl"), "w").write("newmtl wall\nKd 0.7 0.7 0.7\nnewmtl lamp\nKd 0.7 0.7 0.7\nKe 10 10 10\n")
    objects = [{"pos": [0, 0, 0], "ori": [0, 0, 0], "scl": [1, 1, 1], "obj": room, "mtl": os.path.join(tmp, "room.mtl")},
               {"pos": [0, -0.5, 3], "ori": [0, 0, 0], "scl": [2, 2, 2], "obj": sphere, "mtl": "*Kd 0.8 0.6 0.3*"}]
    return objects, n_tris


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--levels", type=int, default=6)
    ap.add_argument("--size", type=int, default=1024)
    ap.add_argument("--spp", type=int, default=16)
    ap.add_argument("--depth", type=int, default=8)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--sort-rays", type=int, default=0, help="sort the incoherent test rays by an n-bit Morton cell + octant first")
    ap.add_argument("--cache-dir", default="", help="keep the generated OBJ files here and reuse them")
    ap.add_argument("--only-render", type=int, default=0, help="1/2: render with that integrator only and exit (profiling); 4: adjoint pass only")
    a = ap.parse_args()
    t0 = time.time()
    objects, n_tris = stress_scene_objects(a.levels, a.cache_dir)
    print("generated %d sphere triangles in %.1f s" % (n_tris, time.time() - t0))
    L = native.lib()
    k0 = L.iptKernelLaunchCount()
    t0 = time.time()
    sc = ipt.Scene(objects)
    t_load = time.time() - t0
    print("loadScene: %.2f s total (OBJ parse + triangle setup + GPU BVH build + upload); nT=%d nE=%d nodes=%d bvh=%d large-triangle set=%d; %d kernel launches" % (
        t_load, sc.nT, sc.nE, sc.n_nodes, sc.uses_bvh, sc.n_large, L.iptKernelLaunchCount() - k0))
    if a.only_render == 3:
        pass
    elif a.only_render == 4:   # adjoint replay (wavefront: recorded vertices + backward sweep), for launch lists
        tgt = np.random.RandomState(1).randint(0, 256, (a.size, a.size, 3)).astype(np.uint8)
        for rep in range(a.reps):
            p = ipt.default_params(width=a.size, height=a.size, spp=a.spp, max_depth=a.depth, seed=rep)
            loss, grad, img = sc.render_grad(p, tgt)
            km = sc.kernel_times_ms(2)
            print("render_grad: loss %.6f |grad|max %.3e  adjoint %.1f ms, forward %.1f ms" % (loss, float(np.abs(grad).max()), km[0], km[1]))
        return
    elif a.only_render:
        sc.set_integrator(a.only_render, 0)
        for rep in range(a.reps):
            p = ipt.default_params(width=a.size, height=a.size, spp=a.spp, max_depth=a.depth, seed=rep)
            img, cnt = sc.render(p)
            ms = sc.kernel_times_ms(1)[0]
            print("render [integrator %d] %s device %.1f ms -> %.1f Mrays/s" % (a.only_render, cnt, ms, (cnt["ext_rays"] + cnt["shadow_rays"]) / ms / 1e3))
        return
    if a.only_render != 3:
        bvh_and_render(a, sc, L)
    trace_test(a, sc, L)


def bvh_and_render(a, sc, L):
    # BVH build alone on the same primitives (host arrays -> CUDA builder -> host arrays)
    tab = sc.triangle_table()
    boxes = np.concatenate([tab["verts"].min(1), tab["verts"].max(1)], 1)
    for rep in range(2):
        t0 = time.time()
        nb, ni, perm = native.build_bvh(boxes, tab["center"], builder=2)
        print("iptBuildBVH (CUDA builder, incl. H2D/D2H of %.0f MB): %.3f s, %d nodes, depth levels ~%d" % (
            (boxes.nbytes + tab["center"].nbytes) / 1e6, time.time() - t0, ni.shape[0], int(np.ceil(np.log2(max(ni.shape[0], 2))))))
    images = {}
    for integ, name in ((1, "megakernel"), (2, "wavefront")):
        sc.set_integrator(integ, 0)
        for rep in range(a.reps):
            p = ipt.default_params(width=a.size, height=a.size, spp=a.spp, max_depth=a.depth, seed=rep)
            img, cnt = sc.render(p)
            ms = sc.kernel_times_ms(1)[0]
            rays = cnt["ext_rays"] + cnt["shadow_rays"]
            print("render [%s] %dx%dx%d depth %d: %s device %.1f ms -> %.1f Mrays/s, %.1f Mpaths/s, mean radiance %.4f" % (
                name, a.size, a.size, a.spp, a.depth, cnt, ms, rays / ms / 1e3, cnt["paths"] / ms / 1e3, float(img.mean())))
            images[(integ, rep)] = img
    sc.set_integrator(0, 0)
    for rep in range(a.reps):
        print("wavefront image == megakernel image (seed %d): %s" % (rep, images[(1, rep)].tobytes() == images[(2, rep)].tobytes()))


def trace_test(a, sc, L):
    for mb in (16, 32, 48, 2048):
        print("device copy of %4d MiB (read + write): %.0f GB/s%s" % (mb, native.measure_copy_gbps(mb << 20, 20 if mb < 1024 else 5),
                                                                  "  <- source + destination fit the 126 MB L2" if 2 * mb <= 120 else ""))
    # incoherent rays only: random origins inside the room, random directions
    rng = np.random.RandomState(0)
    n = 8_000_000
    d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    o = rng.uniform([-1.9, -1.9, 1.1], [1.9, 1.9, 4.9], (n, 3))
    rays = np.concatenate([o, d], 1).astype(np.float32)
    if a.sort_rays:
        # origin cell (Morton order, 2^bits cells per axis) then direction octant: what a device-side ray sort would produce
        bits = a.sort_rays
        q = np.clip(((o - [-2, -2, 1]) / 4.0 * (1 << bits)).astype(np.int64), 0, (1 << bits) - 1)
        key = np.zeros(n, np.int64)
        for b in range(bits):
            for ax in range(3):
                key |= ((q[:, ax] >> b) & 1) << (3 * b + ax)
        key = (key << 3) | ((d[:, 0] > 0).astype(np.int64) | ((d[:, 1] > 0).astype(np.int64) << 1) | ((d[:, 2] > 0).astype(np.int64) << 2))
        rays = rays[np.argsort(key, kind="stable")]
        print("rays sorted by %d-bit Morton cell + octant" % bits)
    import torch
    from inverse_path_tracer_b200.torch_api import TorchScene  # noqa: F401
    import ctypes as C
    rd = torch.from_numpy(rays).cuda()
    tri = torch.empty(n, dtype=torch.int32, device="cuda"); tt = torch.empty(n, dtype=torch.float32, device="cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    stats = torch.zeros(2, dtype=torch.int64, device="cuda")
    L.iptTraceRaysStatsDevice(sc.h, C.c_void_p(rd.data_ptr()), n, C.c_void_p(tri.data_ptr()), C.c_void_p(tt.data_ptr()), C.c_void_p(stats.data_ptr()), st)
    torch.cuda.synchronize()
    boxes_per_ray, tris_per_ray = float(stats[0]) / n, float(stats[1]) / n
    bytes_per_ray = boxes_per_ray * 32 + tris_per_ray * 64
    print("traversal work per incoherent ray: %.1f child boxes (32 B) + %.2f leaf triangles (64 B) = %.0f algorithmic bytes (quantised records: 16 B per box requested)" % (
        boxes_per_ray, tris_per_ray, bytes_per_ray))
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        L.iptTraceRaysDevice(sc.h, C.c_void_p(rd.data_ptr()), n, C.c_void_p(tri.data_ptr()), C.c_void_p(tt.data_ptr()), st)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print("incoherent closest hit: %d rays in %.2f ms -> %.1f Mrays/s (hit rate %.3f); %.0f GB/s of algorithmic node+triangle traffic (HBM copy peak 6555 GB/s)" % (
            n, ms, n / ms / 1e3, float((tri >= 0).float().mean()), n * bytes_per_ray / ms / 1e6))


if __name__ == "__main__":
    main()

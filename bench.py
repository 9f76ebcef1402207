#!/usr/bin/env python
"""bench.py -- the hot path on N B200s of one node, one JSON line on stdout (rank 0).

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference --gpus 1 --steps 2 --warmup 1      # the reference arm (CPU oracle)

Workload (BASELINE.json configs[1]): CornellBox-Empty-CO forward path trace, 1024 x 1024, 1024 spp,
max depth 8, Russian roulette .9, Philox seed = step index.  One step = one full render of the image.
At N > 1 the pixel rows are sharded (interleaved: rank, rank+N, ...), the scene is replicated, and the rows are
all-gathered over NCCL inside the timed region (strong scaling: the image is fixed).

Numbers:
  value        Mrays/s, whole job, device-resident (image left in HBM), CUDA events, max over ranks.
               A ray is one closest-hit query (extension or shadow), counted by device counters.
  e2e          same metric through the host-buffer C-ABI call (iptRender): every step uploads the
               albedos (setMaterials) and reads the float image back to host memory.
  roofline     FP32-pipe roofline of the dominant kernel k_paths<RENDER>: algorithmic FLOPs =
               rays x 36 FLOP x nT triangles over the kernel's own CUDA-event time;
               peak = SMs x 128 lanes x 2 x sm_max_mhz (MEASURED_PEAKS.json).  No tensor cores, and the
               scene lives in shared memory, so neither "hbm" nor "tensor" bounds this kernel.
  cpu_baseline the CPU oracle (oracle/, a port of the reference algorithm -- the reference has no CPU
               renderer) on all host threads, on a bounded row band of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
ASSETS = os.path.join(ROOT, "tests", "golden", "assets")

WORKLOAD = {"name": "CornellBox-Empty-CO forward path trace 1024x1024, 1024 spp, depth 8", "width": 1024, "height": 1024, "spp": 1024,
            "max_depth": 8, "p_rr": 0.9, "group_samples": -1}   # -1: the library's size rule (32 for this job, include/ipt_b200.h)
CORNELL = [{"pos": [0, 0, 4], "ori": [0, 0, 0], "scl": [2, 2, 2], "obj": "./CornellBox/CornellBox-Empty-CO.obj",
            "mtl": "./CornellBox/CornellBox-Empty-CO.mtl"}]
FLOP_PER_TEST = 36   # denom 5 + numerator 6 + div 1 + point 6 + three edge planes 18


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured"
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


class ClockSampler:
    """nvidia-smi sampling during the timed region (B200_PROFILING.md clocks line)"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        rows = [l.strip().split(", ") for l in open(self.f.name) if l.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); power.append(float(r[3]))
            except ValueError:
                continue
            for name, v in zip(names, r[5:9]):
                if v.strip().lower() == "active":
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm),
                "power_w_max": max(power)}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# ----------------------------------------------------------------------------------------------------
# reference arm: the reference's algorithm on the host cores (oracle port; the reference ships no CPU
# renderer and cannot be compiled here -- SURVEY.md R4, 8c)
# ----------------------------------------------------------------------------------------------------
def cpu_band(height, rows):
    b = max(0, height // 2 - rows // 2)
    return b, min(height, b + rows)


def run_cpu_sample(O, osc, seed, rows, threads):
    w = WORKLOAD
    rb, re = cpu_band(w["height"], rows)
    P = O.make_params(w["width"], w["height"], w["spp"], max_depth=w["max_depth"], p_rr=w["p_rr"], seed=seed,
                      group_samples=w["group_samples"], row_begin=rb, row_end=re)
    t0 = time.perf_counter()
    _, cnt = osc.render(P, threads=threads)
    dt = time.perf_counter() - t0
    return cnt, dt, (rb, re)


def cpu_render_rate(O, osc, threads, reps=1, **kw):
    """best of `reps` oracle renders: (Mrays/s, Mpaths/s, seconds, counters)"""
    best = None
    for i in range(reps):
        P = O.make_params(seed=i, **kw)
        t0 = time.perf_counter()
        _, cnt = osc.render(P, threads=threads)
        dt = time.perf_counter() - t0
        if best is None or dt < best[2]:
            best = ((cnt["ext_rays"] + cnt["shadow_rays"]) / dt / 1e6, cnt["paths"] / dt / 1e6, dt, cnt)
    return best


def reference_arm(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    from oracle import oracle as O
    kind = O.use_native_build()
    osc = O.OracleScene(CORNELL, ASSETS)
    threads = O.hardware_threads() or 1
    rows = args.cpu_rows
    for i in range(args.warmup):
        run_cpu_sample(O, osc, 1000 + i, max(1, rows // 4), threads)
    rays = paths = 0
    total = 0.0
    band = None
    for i in range(args.steps):
        cnt, dt, band = run_cpu_sample(O, osc, i, rows, threads)
        rays += cnt["ext_rays"] + cnt["shadow_rays"]
        paths += cnt["paths"]
        total += dt
    v = rays / total / 1e6
    sample = "rows [%d,%d) of the %dx%d image at %d spp per step (%d paths/step), %d host threads, oracle built %s" % (
        band[0], band[1], WORKLOAD["width"], WORKLOAD["height"], WORKLOAD["spp"], paths // max(args.steps, 1), threads, kind)
    line = {"impl": "reference", "metric": "Mrays/s", "value": v, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / max(args.steps, 1), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD["name"], "sampled": sample},
            "cpu_baseline": {"value": v, "unit": "Mrays/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "extra": {"mpaths_per_s": paths / total / 1e6}}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------
# B200 arm
# ----------------------------------------------------------------------------------------------------
def b200_arm(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import inverse_path_tracer_b200 as ipt
    from inverse_path_tracer_b200 import native, parallel
    from inverse_path_tracer_b200.torch_api import TorchScene, inverse_optimizer, render_loss

    rank, world, local = dist_env()
    if world != args.gpus:
        raise SystemExit("bench.py --gpus %d must run under torchrun with WORLD_SIZE=%d (got %d)" % (args.gpus, args.gpus, world))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 arm has no CPU fallback (use --impl reference for the CPU oracle)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    launches0 = native.lib().iptKernelLaunchCount()
    w = dict(WORKLOAD)
    if args.small:   # developer smoke: not a bench number
        w.update(width=256, height=256, spp=64, name=w["name"] + " [SMALL DEV RUN 256x256x64 -- not a bench number]")
    ts = TorchScene(CORNELL, asset_root=ASSETS, device=local)
    nT = ts.nT
    rb, re, rstep = parallel.shard_rows_interleaved(w["height"], rank, world)   # rows rank, rank+N, ...: equal work per GPU
    rows = len(range(rb, re, rstep))
    peaks, peak_kind = measured_peaks()
    sm_count = torch.cuda.get_device_properties(dev).multi_processor_count
    peak_tflops = sm_count * 128 * 2 * peaks["sm_max_mhz"] * 1e6 / 1e12
    peak_source = "%s: %d SMs x 128 lanes x 2 x sm_max_mhz %.0f" % (peak_kind, sm_count, peaks["sm_max_mhz"])

    def fp32_roofline(local_rays_per_launch, n_tris, kernel_ms, traffic=None, traffic_source=None):
        ach = local_rays_per_launch * FLOP_PER_TEST * n_tris / (kernel_ms * 1e-3) / 1e12
        r = {"bound": "fp32", "achieved": ach, "peak": peak_tflops, "unit": "TFLOP/s", "frac": ach / peak_tflops, "traffic": traffic,
             "peak_source": peak_source, "kernel_ms": kernel_ms,
             "algorithmic": "%d FLOP per ray-triangle test x %d triangles per ray (SURVEY.md 8d)" % (FLOP_PER_TEST, n_tris)}
        if traffic_source:
            r["traffic_source"] = traffic_source
        return r

    def params(seed, **kw):
        base = dict(width=w["width"], height=w["height"], spp=w["spp"], max_depth=w["max_depth"], p_rr=w["p_rr"],
                    group_samples=w["group_samples"], row_begin=rb, row_end=re, row_step=rstep, seed=seed)
        base.update(kw)
        return ipt.default_params(**base)

    out = torch.empty((rows, w["width"], 3), dtype=torch.float32, device=dev)
    counters = torch.zeros(3, dtype=torch.int64, device=dev)
    total = torch.zeros(3, dtype=torch.int64, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def step(seed):
        ts.render(params(seed), out=out, counters=counters)
        total.add_(counters)
        img = parallel.gather_rows_interleaved(out, w["height"]) if world > 1 else out
        flush.zero_()
        return img

    # ---- device-resident timing -----------------------------------------------------------------------
    for i in range(args.warmup):
        step(10_000 + i)
    barrier()
    total.zero_()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step(i)
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = max_over_ranks(e0.elapsed_time(e1))
    kernel_ms = ts.scene.kernel_times_ms(args.steps)
    cnt = [int(x) for x in total.tolist()]
    paths = sum_over_ranks(cnt[0])
    rays = sum_over_ranks(cnt[1] + cnt[2])
    value = rays / (ms * 1e-3) / 1e6
    # roofline of the dominant kernel on this rank (rank 0 reports its own)
    local_rays = cnt[1] + cnt[2]
    k_ms = float(np.mean(kernel_ms)) if len(kernel_ms) else float("nan")
    traffic, traffic_source = None, None
    tpath = os.path.join(ROOT, "profiles", "traffic_k_paths_render.json")
    if os.path.exists(tpath) and world == 1 and not args.small:
        try:
            tj = json.load(open(tpath))
            traffic = tj.get("dram_bytes_per_launch")
            traffic_source = "constant from the committed ncu --set full capture of this launch (%s); not re-measured in this run" % tj.get("source", "profiles/")
        except Exception:
            traffic = None
    roofline = fp32_roofline(local_rays / max(args.steps, 1), nT, k_ms, traffic, traffic_source)

    # ---- end to end through the host-buffer C ABI -------------------------------------------------------
    host_img = torch.empty((rows, w["width"], 3), dtype=torch.float32, pin_memory=True).numpy()
    kd = ts.scene.get_materials()
    e2e_steps = max(10, args.steps)

    def e2e_step(seed):
        ts.scene.set_materials(kd)                                      # H2D: this step's inputs (albedos)
        return ts.scene.render(params(seed), img=host_img)[1]           # D2H: the float image of this rank's rows

    if args.small:
        e2e_steps = 3
    e2e_step(20_000)
    barrier()
    t0 = time.perf_counter()
    e_rays = 0
    for i in range(e2e_steps):
        c = e2e_step(i)
        e_rays += c["ext_rays"] + c["shadow_rays"]
    torch.cuda.synchronize()
    e_dt = max_over_ranks(time.perf_counter() - t0)
    e_value = sum_over_ranks(e_rays) / e_dt / 1e6

    extra = {"mpaths_per_s": paths / (ms * 1e-3) / 1e6, "rays_per_path": rays / max(paths, 1),
             "kernel_ms_mean": k_ms, "kernel_share_of_step": k_ms * args.steps / max(ms, 1e-9)}
    if not args.no_extras:
        # ---- BASELINE.json's metric point: CornellBox 512 x 512, 64 spp (unbounded depth: the reference has no depth
        #      limit) -- device-resident value, FP32 roofline of its kernel, end to end through iptRender, CPU oracle.
        #      Each frame is timed by its own event pair; the L2 flush runs between frames, outside the timed intervals.
        hb, he, hs = parallel.shard_rows_interleaved(512, rank, world)
        hp = lambda s: ipt.default_params(width=512, height=512, spp=64, max_depth=0, seed=s, row_begin=hb, row_end=he, row_step=hs)
        hrows = len(range(hb, he, hs))
        hout = torch.empty((hrows, 512, 3), dtype=torch.float32, device=dev)
        for i in range(3):
            ts.render(hp(100 + i), out=hout, counters=counters)
        barrier()
        total.zero_()
        reps = 20
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
        for i in range(reps):
            flush.zero_()
            if world > 1:
                dist.barrier()
            evs[i][0].record()
            ts.render(hp(i), out=hout, counters=counters)
            if world > 1:
                parallel.gather_rows_interleaved(hout, 512)
            evs[i][1].record()
            total.add_(counters)
        barrier()
        hms = max_over_ranks(sum(a.elapsed_time(b) for a, b in evs))
        hk = ts.scene.kernel_times_ms(reps)
        hc = [int(x) for x in total.tolist()]
        h_rays, h_paths = sum_over_ranks(hc[1] + hc[2]), sum_over_ranks(hc[0])
        point = {"mrays_per_s": h_rays / (hms * 1e-3) / 1e6, "mpaths_per_s": h_paths / (hms * 1e-3) / 1e6, "ms_per_frame": hms / reps,
                 "timing": "sum of %d per-frame CUDA-event intervals (render%s), max over ranks; 256 MiB L2 flush between frames, outside the intervals" % (
                     reps, " + all-gather" if world > 1 else ""),
                 "roofline": fp32_roofline((hc[1] + hc[2]) / reps, nT, float(np.mean(hk)))}
        # end to end: host buffers in, host image out
        hhost = torch.empty((hrows, 512, 3), dtype=torch.float32, pin_memory=True).numpy()

        def h_e2e(seed):
            ts.scene.set_materials(kd)
            return ts.scene.render(hp(seed), img=hhost)[1]
        h_e2e(99)
        barrier()
        t0 = time.perf_counter()
        he_rays = 0
        for i in range(20):
            c = h_e2e(i)
            he_rays += c["ext_rays"] + c["shadow_rays"]
        torch.cuda.synchronize()
        he_dt = max_over_ranks(time.perf_counter() - t0)
        point["e2e"] = {"value": sum_over_ranks(he_rays) / he_dt / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": int(nT * 3 * 4),
                        "d2h_bytes_per_step": int(hrows * 512 * 3 * 4 + 24), "steps": 20, "ms_per_frame": 1e3 * he_dt / 20}
        if world == 1:
            # the same frames as a stream of work: two scene handles on two CUDA streams (include/ipt_b200.h "Concurrency": one
            # handle per stream), frames alternating, one event pair around all of them -- the tail of one frame's persistent
            # launch (its longest paths; unbounded depth) overlaps the start of the next frame on the other stream
            ts_b = TorchScene(CORNELL, asset_root=ASSETS, device=local)
            sts = [torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)]
            scs, outs2 = [ts, ts_b], [hout, torch.empty_like(hout)]
            cn2 = [torch.zeros_like(counters), torch.zeros_like(counters)]
            tot2 = [torch.zeros_like(total), torch.zeros_like(total)]
            for k in range(2):
                with torch.cuda.stream(sts[k]):
                    scs[k].render(hp(200 + k), out=outs2[k], counters=cn2[k])
            torch.cuda.synchronize()
            flush.zero_()
            torch.cuda.synchronize()
            ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ea.record()
            for st_ in sts:
                st_.wait_stream(torch.cuda.current_stream())
            for i in range(reps):
                k = i & 1
                with torch.cuda.stream(sts[k]):
                    scs[k].render(hp(i), out=outs2[k], counters=cn2[k])
                    tot2[k].add_(cn2[k])
            for st_ in sts:
                torch.cuda.current_stream().wait_stream(st_)
            eb.record()
            torch.cuda.synchronize()
            ms2 = ea.elapsed_time(eb)
            c2 = [int(x) for x in (tot2[0] + tot2[1]).tolist()]
            point["two_handles_two_streams"] = {"mrays_per_s": (c2[1] + c2[2]) / (ms2 * 1e-3) / 1e6, "mpaths_per_s": c2[0] / (ms2 * 1e-3) / 1e6,
                                                "ms_per_frame": ms2 / reps,
                                                "timing": "one CUDA-event interval around %d frames issued alternately on two scene handles / two streams "
                                                          "(throughput of a stream of frames, not the latency of one; no L2 flush inside: the scene lives in shared memory)" % reps}
            ts_b.close()
        extra["metric_point_cornell_512x512_64spp_unbounded"] = point

        # ---- inverse iterations/s (BASELINE config 3): scenes/0.txt, 500x500 target, 64 spp / iteration, Adam on the 30x3
        #      albedos, rows interleaved over the ranks.  The whole iteration runs inside the library (csrc/inverse.cu:
        #      render, loss, adjoint, GPU-to-GPU exchange of [loss, gradient] over NVLink, Adam) -- one C call per batch.
        from inverse_path_tracer_b200.ipt_cuda import parse_scene_file
        objs = parse_scene_file(os.path.join(ASSETS, "scenes", "0.txt"))
        inv = TorchScene(objs, asset_root=ASSETS, device=local)
        from PIL import Image
        target = torch.from_numpy(np.asarray(Image.open(os.path.join(ROOT, "tests", "golden", "golden_0_true.png")).convert("RGB")).copy()).to(dev)
        ib, ie, istep = parallel.shard_rows_interleaved(500, rank, world)   # 500 rows: shards may differ by one row; only sums are exchanged
        ip = ipt.default_params(width=500, height=500, spp=64, max_depth=0, row_begin=ib, row_end=ie, row_step=istep)
        opt = inverse_optimizer(inv, ip, target, lr=0.02)
        opt.set_kd(np.full((inv.nT, 3), 0.5, np.float32))
        opt.step(1000, 5)
        opt.state()
        barrier()
        iters = 100
        t0 = time.perf_counter()
        opt.step(0, iters)
        loss_after, _, _ = opt.state()
        i_dt = max_over_ranks(time.perf_counter() - t0)
        ik = inv.scene.kernel_times_ms(1)            # forward + loss + adjoint chains of the last iteration, one CUDA-event interval
        ic = inv.scene.last_counters()               # rays of the last adjoint pass (the forward pass traces the same paths)
        i_rays = 2 * (ic["ext_rays"] + ic["shadow_rays"])
        inv_extra = {"iters_per_s": iters / i_dt, "ms_per_iter": 1e3 * i_dt / iters, "loss_after": loss_after, "impl": "iptInverseStep (in-library loop)",
                     "timing": "wall clock around %d iterations enqueued by one call + the final state read-back, max over ranks" % iters,
                     "kernel_ms": float(ik[0]) if len(ik) else None}
        if len(ik):
            inv_extra["roofline"] = fp32_roofline(i_rays, inv.nT, float(ik[0]))
            inv_extra["roofline"]["note"] = "forward + adjoint passes of one iteration on this rank (row chunks on overlapping streams); the adjoint replays the forward paths (same rays)"
        opt.close()
        if world == 1:
            # the Python loop this replaces (render_loss + torch.optim.Adam), for comparison
            kd_param = torch.full((inv.nT, 3), 0.5, device=dev, requires_grad=True)
            adam = torch.optim.Adam([kd_param], lr=0.02)

            def inv_iter(it):
                adam.zero_grad(set_to_none=True)
                loss = render_loss(kd_param, inv, ipt.default_params(width=500, height=500, spp=64, max_depth=0, seed=it), target)
                loss.backward()
                adam.step()
                with torch.no_grad():
                    kd_param.clamp_(1e-3, 1.0)
            for it in range(3):
                inv_iter(1000 + it)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for it in range(20):
                inv_iter(it)
            torch.cuda.synchronize()
            inv_extra["python_loop_iters_per_s"] = 20 / (time.perf_counter() - t0)
        extra["inverse_scene0_500x500_64spp"] = inv_extra

        # ---- scene-sharded batch inverse (BASELINE config 4): every rank fits its own scene (scenes/<rank>.txt) against
        #      its own target (forward render, 100 spp, seed = scene id); no exchange; aggregate iterations/s
        mine = TorchScene(parse_scene_file(os.path.join(ASSETS, "scenes", "%d.txt" % rank)), asset_root=ASSETS, device=local)
        tgt8 = torch.empty((500, 500, 3), dtype=torch.uint8, device=dev)
        mine.render(ipt.default_params(width=500, height=500, spp=100, seed=rank), out_u8=tgt8)
        torch.cuda.synchronize()
        bopt = native.InverseOptimizer(mine.scene, ipt.default_params(width=500, height=500, spp=64, max_depth=0), tgt8.data_ptr(), lr=0.02)
        bopt.set_kd(np.full((mine.nT, 3), 0.5, np.float32))
        bopt.step(1000, 3)
        bopt.state()
        barrier()
        t0 = time.perf_counter()
        bopt.step(0, 40)
        lossb, _, _ = bopt.state()
        b_dt = max_over_ranks(time.perf_counter() - t0)
        extra["inverse_batch_scene_sharded_500x500_64spp"] = {"iters_per_s": world * 40 / b_dt, "scenes_in_flight": world, "ms_per_iter_per_gpu": 1e3 * b_dt / 40,
                                                               "loss_after_rank0": lossb}
        bopt.close()
        mine.close()
        inv.close()

        # ---- BASELINE config 5 (stress): room + shapes/sphere.obj subdivided to 5 242 880 triangles, 1024 x 1024, 16 spp,
        #      depth 8, wavefront integrator; incoherent closest hit with its HBM roofline (algorithmic bytes per ray
        #      from the instrumented traversal, iptTraceRaysStatsDevice).  Rank 0 of a single-GPU run only.
        if world == 1 and not args.no_stress:
            extra["config5_room_sphere_5p2M_triangles"] = config5_extra(args, ipt, native, torch, np, dev, peaks, peak_kind)

    # ---- CPU baseline on this box's host cores (rank 0, N = 1 only) -------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle import oracle as O
        kind = O.use_native_build()
        osc = O.OracleScene(CORNELL, ASSETS)
        threads = O.hardware_threads() or 1
        ccnt, cdt, band = run_cpu_sample(O, osc, 0, args.cpu_rows, threads)
        cpu = {"value": (ccnt["ext_rays"] + ccnt["shadow_rays"]) / cdt / 1e6, "unit": "Mrays/s", "cores": threads, "kind": "port",
               "sample": "rows [%d,%d) of the %dx%d image at %d spp (%d paths, %.1f s), CPU oracle on %d host threads, built %s" % (
                   band[0], band[1], WORKLOAD["width"], WORKLOAD["height"], WORKLOAD["spp"], ccnt["paths"], cdt, threads, kind)}
        if not args.no_extras:
            # the metric point on the CPU: the full 512 x 512 x 64 spp frame
            r, pth, dt, _ = cpu_render_rate(O, osc, threads, width=512, height=512, spp=64, max_depth=0)
            extra["metric_point_cornell_512x512_64spp_unbounded"]["cpu_baseline"] = {
                "value": r, "unit": "Mrays/s", "mpaths_per_s": pth, "cores": threads, "kind": "port",
                "sample": "the whole 512x512x64 spp frame (%.1f s), CPU oracle on %d host threads, built %s" % (dt, threads, kind)}
            # BASELINE config 1 / BASELINE.md section 5: 256 x 256, 16 spp, depth 5, one thread and all threads, best of 3
            c1 = {}
            for name, th in (("one_thread", 1), ("all_threads", threads)):
                r, pth, dt, _ = cpu_render_rate(O, osc, th, reps=3, width=256, height=256, spp=16, max_depth=5)
                c1[name] = {"mrays_per_s": r, "mpaths_per_s": pth, "seconds": dt, "threads": th}
            c1["build"] = kind
            extra["cpu_config1_cornell_256x256_16spp_depth5"] = c1

    launches = native.lib().iptKernelLaunchCount() - launches0
    if rank == 0:
        line = {"metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / max(args.steps, 1), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": w["name"], "scene": "CornellBox-Empty-CO (18 triangles, 2 emissive), scene replicated",
                           "sharding": "interleaved pixel rows (rank, rank+N, ...), all-gather of the float image" if world > 1 else "single GPU",
                           "rng": "Philox4x32-10 keyed (pixel, sample, bounce), seed = step index",
                           "l2": "flushed between iterations by a 256 MiB device write (inside the timed region)",
                           "metric_point": "BASELINE.json quotes its metric on CornellBox 512x512x64 spp: extra.metric_point_cornell_512x512_64spp_unbounded "
                                           "carries that point with its own roofline, e2e and cpu_baseline"},
                "roofline": roofline,
                "cpu_baseline": cpu,
                "e2e": {"value": e_value, "unit": "Mrays/s", "h2d_bytes_per_step": int(nT * 3 * 4),
                        "d2h_bytes_per_step": int(rows * w["width"] * 3 * 4 + 24), "steps": e2e_steps},
                "gpu_launches": int(launches), "clocks": clocks, "extra": extra}
        print(json.dumps(line), flush=True)
    ts.close()
    if world > 1:
        dist.destroy_process_group()


def config5_extra(args, ipt, native, torch, np, dev, peaks, peak_kind):
    import ctypes as C
    import importlib.util
    import tempfile
    spec = importlib.util.spec_from_file_location("stress_config5", os.path.join(ROOT, "tools", "stress_config5.py"))
    stress = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(stress)
    t0 = time.perf_counter()
    sobjs, sphere_tris = stress.stress_scene_objects(6, tempfile.mkdtemp(prefix="ipt_bench_"))
    t_gen = time.perf_counter() - t0
    L = native.lib()
    k0 = L.iptKernelLaunchCount()
    t0 = time.perf_counter()
    sc = ipt.Scene(sobjs)
    t_load = time.perf_counter() - t0
    tab = sc.triangle_table()
    boxes = np.concatenate([tab["verts"].min(1), tab["verts"].max(1)], 1)
    t0 = time.perf_counter()
    nb, ni, perm = native.build_bvh(boxes, tab["center"], builder=2)
    t_bvh = time.perf_counter() - t0
    out = {"triangles": int(sc.nT), "bvh_nodes": int(sc.n_nodes), "large_triangle_set": int(sc.n_large), "generate_obj_s": t_gen,
           "load_scene_s": t_load, "bvh_build_s": t_bvh,
           "bvh_build_note": "iptBuildBVH, CUDA builder, %d primitives, including the host<->device copies" % boxes.shape[0]}
    # path tracing through the wavefront integrator
    p = lambda s_: ipt.default_params(width=1024, height=1024, spp=16, max_depth=8, seed=s_)
    sc.render(p(50))
    best = None
    for i in range(3):
        _, cnt = sc.render(p(i))
        ms = float(sc.kernel_times_ms(1)[0])
        if best is None or ms < best[0]:
            best = (ms, cnt)
    out["path_tracing_1024x1024_16spp_depth8"] = {"integrator": "wavefront", "ms_per_frame": best[0], "mrays_per_s": (best[1]["ext_rays"] + best[1]["shadow_rays"]) / best[0] / 1e3,
                                                  "mpaths_per_s": best[1]["paths"] / best[0] / 1e3}
    # adjoint replay of the same frame through the wavefront (vertex record + backward sweep), depth 8
    tgt = np.random.RandomState(1).randint(0, 256, (1024, 1024, 3)).astype(np.uint8)
    sc.render_grad(p(60), tgt)
    best_adj = None
    for i in range(2):
        sc.render_grad(p(i), tgt)
        km = sc.kernel_times_ms(2)      # newest first: adjoint replay, forward pass
        if best_adj is None or km[0] < best_adj[0]:
            best_adj = (float(km[0]), float(km[1]))
    out["adjoint_1024x1024_16spp_depth8"] = {"integrator": "wavefront (recorded vertices, k_wf_sweep)", "adjoint_ms": best_adj[0], "forward_ms": best_adj[1]}
    # incoherent closest hit: random origins inside the room, random directions
    rng = np.random.RandomState(0)
    n = 8_000_000
    d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    o = rng.uniform([-1.9, -1.9, 1.1], [1.9, 1.9, 4.9], (n, 3))
    rd = torch.from_numpy(np.concatenate([o, d], 1).astype(np.float32)).to(dev)
    tri = torch.empty(n, dtype=torch.int32, device=dev); tt = torch.empty(n, dtype=torch.float32, device=dev)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    stats = torch.zeros(2, dtype=torch.int64, device=dev)
    L.iptTraceRaysStatsDevice(sc.h, C.c_void_p(rd.data_ptr()), n, C.c_void_p(tri.data_ptr()), C.c_void_p(tt.data_ptr()), C.c_void_p(stats.data_ptr()), st)
    torch.cuda.synchronize()
    boxes_per_ray, tris_per_ray = float(stats[0]) / n, float(stats[1]) / n
    bytes_per_ray = boxes_per_ray * 32 + tris_per_ray * 64
    best_ms = None
    for rep in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        L.iptTraceRaysDevice(sc.h, C.c_void_p(rd.data_ptr()), n, C.c_void_p(tri.data_ptr()), C.c_void_p(tt.data_ptr()), st)
        e1.record(); torch.cuda.synchronize()
        m = e0.elapsed_time(e1)
        best_ms = m if best_ms is None or m < best_ms else best_ms
    ach = n * bytes_per_ray / (best_ms * 1e-3) / 1e9
    out["incoherent_closest_hit"] = {"rays": n, "ms": best_ms, "mrays_per_s": n / best_ms / 1e3, "hit_rate": float((tri >= 0).float().mean()),
                                     "child_boxes_per_ray": boxes_per_ray, "leaf_triangles_per_ray": tris_per_ray,
                                     "roofline": {"bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": ach / peaks["hbm_gbs"], "traffic": None,
                                                  "peak_source": "%s HBM copy bandwidth (MEASURED_PEAKS.json)" % peak_kind,
                                                  "l2_resident_copy_gbs": native.measure_copy_gbps(32 << 20, 20),
                                                  "algorithmic": "%.1f child boxes x 32 B + %.2f leaf triangles x 64 B = %.0f bytes per ray (iptTraceRaysStatsDevice)" % (
                                                      boxes_per_ray, tris_per_ray, bytes_per_ray),
                                                  "note": "algorithmic = the reference's data structure (32-byte boxes, SURVEY.md 8d), as in round 1; the shipped traversal records "
                                                          "hold a child box in 16 bytes and a leaf visit reads 16 B planes first, so the bytes actually requested are about "
                                                          "%.0f per ray and `frac` overstates the share of the HBM peak this kernel moves" % (boxes_per_ray * 16 + tris_per_ray * 32)}}
    out["kernel_launches"] = int(L.iptKernelLaunchCount() - k0)
    sc.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-rows", type=int, default=128, help="rows of the 1024-wide image rendered per CPU sample (128 rows = 134 M paths, about 13 s on 16 host threads)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--no-stress", action="store_true", help="skip the 5.2 M-triangle config-5 extra (about 40 s)")
    ap.add_argument("--small", action="store_true", help="developer smoke run (not a bench number)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200" and not args.small:
        args.warmup = 3   # timing rule: at least three warm-up steps
    if args.impl == "reference":
        reference_arm(args)
        return
    rank, world, _ = dist_env()
    if args.gpus > 1 and "RANK" not in os.environ:
        # convenience: re-launch under torchrun on this node
        port = 29500 + (os.getpid() % 2000)
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus), "--master-addr", "127.0.0.1",
               "--master-port", str(port), os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    b200_arm(args)


if __name__ == "__main__":
    main()

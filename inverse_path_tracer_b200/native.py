"""ctypes binding of libipt_b200.so (C ABI: include/ipt_b200.h).

There is no CPU fallback: importing works without a GPU (so the symbol table can be checked), but
any compute entry point fails with the library's own error when no CUDA device is present, and
`lib()` raises if the shared library has not been built (python -m inverse_path_tracer_b200.build).
"""
import ctypes as C
import os

import numpy as np

PKG = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.environ.get("IPT_B200_LIB") or os.path.join(PKG, "lib", "libipt_b200.so")   # override: developer A/B builds


class IptRenderParams(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("spp", C.c_int), ("max_depth", C.c_int),
                ("p_rr", C.c_float), ("seed_lo", C.c_uint32), ("seed_hi", C.c_uint32),
                ("group_samples", C.c_int), ("row_begin", C.c_int), ("row_end", C.c_int), ("row_step", C.c_int)]


class IptInverseConfig(C.Structure):
    _fields_ = [("lr", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float), ("kd_min", C.c_float), ("kd_max", C.c_float),
                ("rank", C.c_int), ("world", C.c_int), ("timeout_ms", C.c_int)]


def library_path():
    return _LIB_PATH


_lib = None

# every symbol include/ipt_b200.h declares
EXPORTS = ["loadScene", "freeScene", "createImage", "createGraph", "getMaterials", "setMaterials",
           "iptGetLastError", "iptDefaultParams", "iptSetAssetRoot", "iptSetDevice", "iptSetTraversalMode",
           "iptSceneInfo", "iptGetTriangleTable", "iptGetCamera", "iptTraceRays", "iptTraceRaysDevice",
           "iptRender", "iptRenderDevice", "iptCreateImageEx", "iptCreateGraphEx", "iptGraphAccumulateDevice",
           "iptGraphCompress", "iptRenderGrad", "iptRenderGradDevice", "iptSetMaterialsDevice", "iptBuildBVH",
           "iptWritePNG", "iptReadPNG", "iptKernelLaunchCount", "iptPathKernelTimes", "iptLastCounters", "iptSetBvhBuilder", "iptTraceRaysStatsDevice",
           "iptSetIntegrator", "iptSceneSetIntegrator", "iptMeasureCopyGBps",
           "iptSetBvhLargeTriangleLimit", "iptSceneLargeTriangles", "iptGraphCompressDevice",
           "iptInverseDefaultConfig", "iptInverseCreate", "iptInverseMailbox", "iptInverseConnect", "iptInverseSetKd", "iptInverseStep",
           "iptInverseState", "iptInverseDestroy"]


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise RuntimeError("libipt_b200.so is not built (%s missing); run `python -m inverse_path_tracer_b200.build`. "
                               "There is no CPU fallback." % _LIB_PATH)
        L = C.CDLL(_LIB_PATH)
        L.iptGetLastError.restype = C.c_char_p
        L.iptKernelLaunchCount.restype = C.c_ulonglong
        L.iptMeasureCopyGBps.restype = C.c_double
        L.iptMeasureCopyGBps.argtypes = [C.c_longlong, C.c_int]
        L.loadScene.restype = C.c_int
        vp = C.c_void_p
        L.freeScene.argtypes = [vp]
        L.getMaterials.argtypes = [vp, vp]
        L.setMaterials.argtypes = [vp, vp]
        L.createImage.argtypes = [vp, C.c_char_p]
        L.createGraph.argtypes = [vp, C.c_char_p, vp]
        L.iptSceneInfo.argtypes = [vp] + [vp] * 5
        L.iptGetTriangleTable.argtypes = [vp] * 10
        L.iptGetCamera.argtypes = [vp, vp]
        L.iptTraceRays.argtypes = [vp, vp, C.c_int, vp, vp]
        L.iptTraceRaysDevice.argtypes = [vp, vp, C.c_int, vp, vp, vp]
        L.iptTraceRaysStatsDevice.argtypes = [vp, vp, C.c_int, vp, vp, vp, vp]
        L.iptRender.argtypes = [vp, vp, vp, vp, vp]
        L.iptRenderDevice.argtypes = [vp, vp, vp, vp, vp, vp]
        L.iptCreateImageEx.argtypes = [vp, vp, C.c_char_p]
        L.iptCreateGraphEx.argtypes = [vp, vp, vp, vp, vp, vp]
        L.iptGraphAccumulateDevice.argtypes = [vp, vp, vp, vp, vp]
        L.iptGraphCompress.argtypes = [vp, C.c_int, vp]
        L.iptRenderGrad.argtypes = [vp, vp, vp, vp, vp, vp]
        L.iptRenderGradDevice.argtypes = [vp, vp, vp, vp, vp, vp]
        L.iptSetMaterialsDevice.argtypes = [vp, vp, vp]
        L.iptBuildBVH.argtypes = [vp, vp, C.c_int, C.c_int, vp, vp, C.c_int, vp]
        L.iptWritePNG.argtypes = [C.c_char_p, vp, C.c_int, C.c_int]
        L.iptReadPNG.argtypes = [C.c_char_p, vp, C.c_int, vp, vp]
        L.iptSetAssetRoot.argtypes = [C.c_char_p]
        L.iptPathKernelTimes.argtypes = [vp, vp, C.c_int]
        L.iptLastCounters.argtypes = [vp, vp]
        L.iptInverseCreate.argtypes = [vp, vp, vp, vp, vp]
        L.iptInverseMailbox.argtypes = [vp, vp, vp]
        L.iptInverseConnect.argtypes = [vp, vp, vp]
        L.iptInverseSetKd.argtypes = [vp, vp, C.c_int]
        L.iptInverseStep.argtypes = [vp, C.c_ulonglong, C.c_int]
        L.iptInverseState.argtypes = [vp, vp, vp, vp, vp, C.c_int]
        L.iptInverseDestroy.argtypes = [vp]
        _lib = L
    return _lib


def last_error():
    return lib().iptGetLastError().decode()


def _check(rc, what):
    if rc != 0:
        raise RuntimeError("%s failed: %s" % (what, last_error()))


def default_params(**kw):
    p = IptRenderParams()
    lib().iptDefaultParams(C.byref(p))
    seed = kw.pop("seed", None)
    if seed is not None:
        p.seed_lo = seed & 0xFFFFFFFF
        p.seed_hi = (seed >> 32) & 0xFFFFFFFF
    for k, v in kw.items():
        if not hasattr(p, k):
            raise TypeError("unknown render parameter %r" % k)
        setattr(p, k, v)
    return p


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def measure_copy_gbps(nbytes, reps=20):
    """read + write GB/s of a device-to-device copy of nbytes (L2-resident when 2 * nbytes fits the L2)"""
    v = lib().iptMeasureCopyGBps(int(nbytes), int(reps))
    if v < 0:
        raise RuntimeError("iptMeasureCopyGBps failed: " + last_error())
    return v


def set_bvh_large_triangle_limit(max_count=128):
    """cap of the brute-force set of large triangles in BVH scenes loaded afterwards (0: all triangles in the tree)"""
    if lib().iptSetBvhLargeTriangleLimit(int(max_count)) != 0:
        raise ValueError(last_error())


def set_integrator(mode=0, wavefront_slots=0):
    """Forward integrator of BVH scenes LOADED AFTERWARDS from this thread: 0 auto (wavefront), 1 persistent
    megakernel, 2 wavefront; wavefront_slots bounds the paths in flight per batch (0 = default).
    Scene.set_integrator changes one loaded scene."""
    if lib().iptSetIntegrator(int(mode), int(wavefront_slots)) != 0:
        raise ValueError(last_error())


def build_bvh(boxes, pos, leaf_size=4, builder=0):
    """BVH::build (bvh.h:109-205) on caller primitives -> (node_boxes, node_ints, perm).
    builder: 0 auto (CUDA kernels from 2048 primitives on), 1 host, 2 CUDA kernels."""
    lib().iptSetBvhBuilder(int(builder))
    boxes = np.ascontiguousarray(boxes, np.float32).reshape(-1, 6)
    pos = np.ascontiguousarray(pos, np.float32).reshape(-1, 3)
    n = boxes.shape[0]
    mx = max(1, 2 * n)
    nb = np.zeros((mx, 6), np.float32)
    ni = np.zeros((mx, 3), np.int32)
    perm = np.zeros(max(n, 1), np.int32)
    k = lib().iptBuildBVH(_ptr(boxes), _ptr(pos), n, leaf_size, _ptr(nb), _ptr(ni), mx, _ptr(perm))
    lib().iptSetBvhBuilder(0)
    if k < 0:
        raise RuntimeError("iptBuildBVH failed: " + last_error())
    return nb[:k], ni[:k], perm[:n]


class Scene:
    """A loaded scene (opaque handle of loadScene, scene.h:177-193)."""

    def __init__(self, objects, asset_root=None, device=None, traversal_mode=0):
        L = lib()
        n = len(objects)
        self._keep = []
        f3 = C.c_float * 3
        fp = C.POINTER(C.c_float)

        def col(key):
            arrs = [f3(*[float(x) for x in o[key]]) for o in objects]
            self._keep.append(arrs)
            return (fp * n)(*[C.cast(a, fp) for a in arrs])

        poss, oris, scls = col("pos"), col("ori"), col("scl")
        objs = (C.c_char_p * n)(*[o["obj"].encode() for o in objects])
        mtls = (C.c_char_p * n)(*[o["mtl"].encode() for o in objects])
        L.iptSetAssetRoot((asset_root or "").encode())
        L.iptSetDevice(-1 if device is None else int(device))
        L.iptSetTraversalMode(int(traversal_mode))
        h = C.c_void_p(0)
        self.nT = L.loadScene(poss, oris, scls, objs, mtls, n, C.byref(h))
        L.iptSetTraversalMode(0)
        if not h.value:
            raise RuntimeError("loadScene failed: " + last_error())
        self.h = h
        info = (C.c_int * 5)()
        ip = [C.cast(C.byref(info, 4 * i), C.c_void_p) for i in range(5)]
        _check(L.iptSceneInfo(self.h, *ip), "iptSceneInfo")
        self.nT, self.nE, self.nO, self.n_nodes, self.uses_bvh = [int(x) for x in info]
        L.iptSceneLargeTriangles.argtypes = [C.c_void_p]
        self.n_large = int(L.iptSceneLargeTriangles(self.h))   # BVH scenes: triangles kept out of the tree

    @classmethod
    def from_file(cls, scenefile, asset_root=None, **kw):
        from .ipt_cuda import parse_scene_file
        return cls(parse_scene_file(scenefile), asset_root=asset_root, **kw)

    def set_integrator(self, mode=0, wavefront_slots=0):
        """forward integrator of this (BVH) scene: 0 auto (wavefront), 1 persistent megakernel, 2 wavefront"""
        L = lib()
        L.iptSceneSetIntegrator.argtypes = [C.c_void_p, C.c_int, C.c_int]
        _check(L.iptSceneSetIntegrator(self.h, int(mode), int(wavefront_slots)), "iptSceneSetIntegrator")

    def close(self):
        if getattr(self, "h", None) is not None and self.h.value:
            lib().freeScene(self.h)
            self.h = C.c_void_p(0)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- tables -------------------------------------------------------------------------------
    def triangle_table(self):
        nT = self.nT
        t = {"verts": np.zeros((nT, 3, 3), np.float32), "kd": np.zeros((nT, 3), np.float32),
             "ks": np.zeros((nT, 3), np.float32), "ke": np.zeros((nT, 3), np.float32),
             "ns": np.zeros(nT, np.float32), "has_vn": np.zeros(nT, np.int32),
             "normal": np.zeros((nT, 3), np.float32), "center": np.zeros((nT, 3), np.float32),
             "area": np.zeros(nT, np.float32)}
        _check(lib().iptGetTriangleTable(self.h, _ptr(t["verts"]), _ptr(t["kd"]), _ptr(t["ks"]), _ptr(t["ke"]), _ptr(t["ns"]),
                                         _ptr(t["has_vn"]), _ptr(t["normal"]), _ptr(t["center"]), _ptr(t["area"])),
               "iptGetTriangleTable")
        return t

    def camera(self):
        M = np.zeros((4, 4), np.float32)
        _check(lib().iptGetCamera(self.h, _ptr(M)), "iptGetCamera")
        return M

    def get_materials(self):
        m = np.zeros((self.nT, 3), np.float32)
        lib().getMaterials(self.h, _ptr(m))
        return m

    def set_materials(self, m):
        m = np.ascontiguousarray(m, np.float32).reshape(self.nT, 3)
        lib().setMaterials(self.h, _ptr(m))

    # ---- hot path, host buffers -----------------------------------------------------------------
    def trace(self, rays):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 6)
        n = rays.shape[0]
        tri = np.zeros(n, np.int32)
        t = np.zeros(n, np.float32)
        _check(lib().iptTraceRays(self.h, _ptr(rays), n, _ptr(tri), _ptr(t)), "iptTraceRays")
        return tri, t

    @staticmethod
    def _rows(p):
        re = p.row_end if 0 < p.row_end <= p.height else p.height
        step = max(p.row_step, 1)
        return (re - max(p.row_begin, 0) + step - 1) // step

    def render(self, p, want_u8=False, img=None, img8=None):
        rows = self._rows(p)
        if img is None:
            img = np.zeros((rows, p.width, 3), np.float32)
        if want_u8 and img8 is None:
            img8 = np.zeros((rows, p.width, 3), np.uint8)
        cnt = np.zeros(3, np.uint64)
        _check(lib().iptRender(self.h, C.byref(p), _ptr(img), _ptr(img8) if want_u8 else None, _ptr(cnt)), "iptRender")
        counters = {"paths": int(cnt[0]), "ext_rays": int(cnt[1]), "shadow_rays": int(cnt[2])}
        return (img, img8, counters) if want_u8 else (img, counters)

    def create_image(self, p, path):
        _check(lib().iptCreateImageEx(self.h, C.byref(p), path.encode()), "iptCreateImageEx")

    def create_graph(self, p, target_u8):
        nT = self.nT
        target = np.ascontiguousarray(target_u8, np.uint8)
        assert target.shape == (p.height, p.width, 3)
        raw = np.zeros(((nT + 1) * nT, 16), np.float64)
        data = np.zeros(7 * (nT + 1) * nT, np.float32)
        cnt = np.zeros(3, np.uint64)
        _check(lib().iptCreateGraphEx(self.h, C.byref(p), _ptr(target), _ptr(raw), _ptr(data), _ptr(cnt)), "iptCreateGraphEx")
        return raw, data, {"paths": int(cnt[0]), "ext_rays": int(cnt[1]), "shadow_rays": int(cnt[2])}

    def kernel_times_ms(self, count=1):
        """device times of the most recent path-kernel launches, newest first"""
        ms = np.zeros(count, np.float32)
        k = lib().iptPathKernelTimes(self.h, _ptr(ms), count)
        if k < 0:
            raise RuntimeError("iptPathKernelTimes failed: " + last_error())
        return ms[:k]

    def last_counters(self):
        cnt = np.zeros(3, np.uint64)
        _check(lib().iptLastCounters(self.h, _ptr(cnt)), "iptLastCounters")
        return {"paths": int(cnt[0]), "ext_rays": int(cnt[1]), "shadow_rays": int(cnt[2])}

    def render_grad(self, p, target_u8):
        target = np.ascontiguousarray(target_u8, np.uint8)
        assert target.shape == (p.height, p.width, 3)
        loss = C.c_double(0)
        grad = np.zeros((self.nT, 3), np.float64)
        img = np.zeros((self._rows(p), p.width, 3), np.float32)
        _check(lib().iptRenderGrad(self.h, C.byref(p), _ptr(target), C.byref(loss), _ptr(grad), _ptr(img)), "iptRenderGrad")
        return loss.value, grad, img


def graph_compress(raw, nT):
    raw = np.ascontiguousarray(raw, np.float64)
    data = np.zeros(7 * (nT + 1) * nT, np.float32)
    _check(lib().iptGraphCompress(_ptr(raw), nT, _ptr(data)), "iptGraphCompress")
    return data


class InverseOptimizer:
    """The inverse loop inside the library (include/ipt_b200.h part 3): Adam on the scene's nT x 3 albedos against a
    target image, one call per iteration (or per batch of iterations), on this GPU's row shard.

    `target_dev_ptr`: device address of the height x width x 3 uint8 target on the scene's GPU (the caller keeps the
    buffer alive).  world > 1: exchange `ipc_handle()` between the ranks and call `connect_ipc(handles)` (one process
    per GPU) or `connect_local([mailbox pointers])` (several ranks in one process)."""

    def __init__(self, scene, params, target_dev_ptr, rank=0, world=1, lr=0.02, beta1=0.9, beta2=0.999, eps=1e-8, kd_min=1e-3, kd_max=1.0, timeout_ms=10000):
        L = lib()
        cfg = IptInverseConfig()
        L.iptInverseDefaultConfig(C.byref(cfg))
        cfg.lr, cfg.beta1, cfg.beta2, cfg.eps, cfg.kd_min, cfg.kd_max = lr, beta1, beta2, eps, kd_min, kd_max
        cfg.rank, cfg.world, cfg.timeout_ms = int(rank), int(world), int(timeout_ms)
        self.scene, self.nT, self.rank, self.world = scene, scene.nT, int(rank), int(world)
        h = C.c_void_p(0)
        _check(L.iptInverseCreate(scene.h, C.byref(params), C.c_void_p(int(target_dev_ptr)), C.byref(cfg), C.byref(h)), "iptInverseCreate")
        self.h = h

    def mailbox(self):
        p = C.c_void_p(0)
        _check(lib().iptInverseMailbox(self.h, C.byref(p), None), "iptInverseMailbox")
        return p.value

    def ipc_handle(self):
        buf = (C.c_ubyte * 64)()
        _check(lib().iptInverseMailbox(self.h, None, buf), "iptInverseMailbox")
        return bytes(buf)

    def connect_local(self, mailboxes):
        arr = (C.c_void_p * self.world)(*[C.c_void_p(m) for m in mailboxes])
        _check(lib().iptInverseConnect(self.h, arr, None), "iptInverseConnect")

    def connect_ipc(self, handles):
        blob = b"".join(handles)
        assert len(blob) == 64 * self.world
        _check(lib().iptInverseConnect(self.h, None, blob), "iptInverseConnect")

    def set_kd(self, kd, reset_moments=True):
        kd = np.ascontiguousarray(kd, np.float32).reshape(self.nT, 3)
        _check(lib().iptInverseSetKd(self.h, _ptr(kd), 1 if reset_moments else 0), "iptInverseSetKd")

    def step(self, first_seed, n_steps=1):
        """enqueue n_steps iterations (seeds first_seed, first_seed + 1, ...); asynchronous"""
        _check(lib().iptInverseStep(self.h, int(first_seed), int(n_steps)), "iptInverseStep")

    def state(self, history=0):
        """wait, then (loss of the last step, albedos [nT, 3], steps applied[, the last `history` losses])"""
        loss = C.c_double(0)
        steps = C.c_ulonglong(0)
        kd = np.zeros((self.nT, 3), np.float32)
        hist = np.zeros(max(history, 1), np.float64)
        _check(lib().iptInverseState(self.h, C.byref(loss), _ptr(kd), C.byref(steps), _ptr(hist) if history else None, int(history)), "iptInverseState")
        return (loss.value, kd, int(steps.value), hist[:history]) if history else (loss.value, kd, int(steps.value))

    def close(self):
        if getattr(self, "h", None) is not None and self.h.value:
            lib().iptInverseDestroy(self.h)
            self.h = C.c_void_p(0)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

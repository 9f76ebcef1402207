"""ctypes front-end of the CPU oracle (oracle/libipt_oracle.so).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; the product package never does.  Build the library with `make -C oracle`.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libipt_oracle.so")
REF_DUMP = os.path.join(HERE, "_ref", "ref_loader_dump")


class OrcParams(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("spp", C.c_int), ("max_depth", C.c_int),
                ("p_rr", C.c_float), ("seed_lo", C.c_uint32), ("seed_hi", C.c_uint32),
                ("group_samples", C.c_int), ("row_begin", C.c_int), ("row_end", C.c_int), ("row_step", C.c_int)]


def make_params(width=500, height=500, spp=100, max_depth=0, p_rr=0.9, seed=0, group_samples=-1,
                row_begin=0, row_end=0, row_step=1):
    return OrcParams(width, height, spp, max_depth, p_rr, seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF,
                     group_samples, row_begin, row_end, row_step)


def n_rows(params):
    re = params.row_end if 0 < params.row_end <= params.height else params.height
    step = max(params.row_step, 1)
    return (re - max(params.row_begin, 0) + step - 1) // step


def build(force=False):
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(
            os.path.join(HERE, "ipt_oracle.cpp")):
        subprocess.check_call(["make", "-C", HERE, "libipt_oracle.so"], stdout=subprocess.DEVNULL)
    if os.path.isdir("/root/reference") and not os.path.exists(REF_DUMP):
        subprocess.check_call(["make", "-C", HERE, "ref"], stdout=subprocess.DEVNULL)


_lib = None
_lib_kind = "-O2 -march=x86-64-v3 (portable build, oracle/Makefile)"
NATIVE_DIR = os.path.join(HERE, "_native")


def use_native_build():
    """Timing legs of bench.py only: rebuild the oracle ON THIS HOST with -O3 -march=native into
    oracle/_native/ (git- and gpurun-ignored: a -march=native binary must not travel between machines) and load that
    one.  Same source, -ffp-contract=off, so the results are the same bits.  Falls back to the portable build when no
    compiler is at hand.  Must be called before the first lib() / OracleScene."""
    global LIB_PATH, _lib_kind
    if _lib is not None:
        return _lib_kind
    out = os.path.join(NATIVE_DIR, "libipt_oracle_native.so")
    src = os.path.join(HERE, "ipt_oracle.cpp")
    try:
        if not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src):
            os.makedirs(NATIVE_DIR, exist_ok=True)
            subprocess.check_call([os.environ.get("CXX", "g++"), "-O3", "-march=native", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-fPIC",
                                   "-pthread", "-shared", "-o", out, src], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        LIB_PATH = out
        _lib_kind = "-O3 -march=native, built on this host"
    except (OSError, subprocess.CalledProcessError):
        pass
    return _lib_kind


def build_kind():
    return _lib_kind


def lib():
    global _lib
    if _lib is None:
        if LIB_PATH.startswith(NATIVE_DIR):
            pass
        else:
            build()
        L = C.CDLL(LIB_PATH)
        L.orc_load_scene.restype = C.c_void_p
        L.orc_load_scene.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_char_p]
        L.orc_last_error.restype = C.c_char_p
        for name in ("orc_free_scene", "orc_num_triangles", "orc_num_emissives", "orc_num_objects"):
            getattr(L, name).argtypes = [C.c_void_p]
        _lib = L
    return _lib


def parse_scene_file(path):
    """Scene-file grammar of the reference (ipt_cuda.py:39-59, 91-107): blocks start with OBJECT;
    POS/ORI/SCL x y z, OBJ path, MTL path | *Kd r g b*; defaults 0/0/1."""
    objs, cur = [], None
    for line in open(path):
        line = line.strip()
        if line == "OBJECT":
            cur = {"pos": [0.0] * 3, "ori": [0.0] * 3, "scl": [1.0] * 3, "obj": None, "mtl": None}
            objs.append(cur)
            continue
        if not line or cur is None:
            continue
        items = line.split(" ")
        tok, vals = items[0], items[1:]
        if tok == "POS":
            cur["pos"] = [float(x) for x in vals]
        elif tok == "ORI":
            cur["ori"] = [float(x) for x in vals]
        elif tok == "SCL":
            cur["scl"] = [float(x) for x in vals]
        elif tok == "OBJ":
            cur["obj"] = vals[0]
        elif tok == "MTL":
            cur["mtl"] = " ".join(vals)
    return objs


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class OracleScene:
    def __init__(self, objects, root=None):
        L = lib()
        n = len(objects)
        pos = np.array([o["pos"] for o in objects], np.float32).reshape(n, 3)
        ori = np.array([o["ori"] for o in objects], np.float32).reshape(n, 3)
        scl = np.array([o["scl"] for o in objects], np.float32).reshape(n, 3)
        objs = (C.c_char_p * n)(*[o["obj"].encode() for o in objects])
        mtls = (C.c_char_p * n)(*[o["mtl"].encode() for o in objects])
        self.h = L.orc_load_scene(_ptr(pos), _ptr(ori), _ptr(scl), objs, mtls, n, (root or "").encode())
        if not self.h:
            raise RuntimeError("oracle load failed: " + L.orc_last_error().decode())
        self.nT = L.orc_num_triangles(self.h)
        self.nE = L.orc_num_emissives(self.h)
        self.nO = L.orc_num_objects(self.h)

    @classmethod
    def from_file(cls, scenefile, root=None):
        return cls(parse_scene_file(scenefile), root)

    def close(self):
        if self.h:
            lib().orc_free_scene(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def triangle_table(self):
        nT = self.nT
        t = {"verts": np.zeros((nT, 3, 3), np.float32), "kd": np.zeros((nT, 3), np.float32),
             "ks": np.zeros((nT, 3), np.float32), "ke": np.zeros((nT, 3), np.float32),
             "ns": np.zeros(nT, np.float32), "has_vn": np.zeros(nT, np.int32),
             "normal": np.zeros((nT, 3), np.float32), "center": np.zeros((nT, 3), np.float32),
             "area": np.zeros(nT, np.float32), "vnormals": np.zeros((nT, 3, 3), np.float32)}
        lib().orc_get_triangle_table(C.c_void_p(self.h), _ptr(t["verts"]), _ptr(t["kd"]), _ptr(t["ks"]), _ptr(t["ke"]),
                                     _ptr(t["ns"]), _ptr(t["has_vn"]), _ptr(t["normal"]), _ptr(t["center"]),
                                     _ptr(t["area"]), _ptr(t["vnormals"]))
        return t

    def get_materials(self):
        m = np.zeros((self.nT, 3), np.float32)
        lib().orc_get_materials(C.c_void_p(self.h), _ptr(m))
        return m

    def set_materials(self, m):
        m = np.ascontiguousarray(m, np.float32).reshape(self.nT, 3)
        lib().orc_set_materials(C.c_void_p(self.h), _ptr(m))

    def camera(self):
        M = np.zeros((4, 4), np.float32)
        lib().orc_get_camera(C.c_void_p(self.h), _ptr(M))
        return M

    def trace(self, rays, threads=8):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 6)
        n = rays.shape[0]
        tri = np.zeros(n, np.int32)
        t = np.zeros(n, np.float32)
        lib().orc_trace(C.c_void_p(self.h), _ptr(rays), n, _ptr(tri), _ptr(t), threads)
        return tri, t

    def render(self, params, threads=8, want_u8=False):
        rows = n_rows(params)
        img = np.zeros((rows, params.width, 3), np.float32)
        img8 = np.zeros((rows, params.width, 3), np.uint8) if want_u8 else None
        cnt = np.zeros(3, np.uint64)
        lib().orc_render(C.c_void_p(self.h), C.byref(params), _ptr(img), _ptr(img8) if want_u8 else None, _ptr(cnt), threads)
        counters = {"paths": int(cnt[0]), "ext_rays": int(cnt[1]), "shadow_rays": int(cnt[2])}
        return (img, img8, counters) if want_u8 else (img, counters)

    def sample(self, params, r, c, s, max_tris=256):
        L = np.zeros(3, np.float32)
        tris = np.full(max_tris, -2, np.int32)
        n = lib().orc_sample(C.c_void_p(self.h), C.byref(params), r, c, s, _ptr(L), _ptr(tris), max_tris)
        return L, tris[:min(n, max_tris)]

    def create_graph(self, params, target_u8, threads=8):
        nT = self.nT
        target = np.ascontiguousarray(target_u8, np.uint8)
        assert target.shape == (params.height, params.width, 3)
        raw = np.zeros(((nT + 1) * nT, 16), np.float64)
        data = np.zeros(7 * (nT + 1) * nT, np.float32)
        cnt = np.zeros(3, np.uint64)
        lib().orc_create_graph(C.c_void_p(self.h), C.byref(params), _ptr(target), _ptr(raw), _ptr(data), _ptr(cnt), threads)
        return raw, data, {"paths": int(cnt[0]), "ext_rays": int(cnt[1]), "shadow_rays": int(cnt[2])}

    def render_grad(self, params, target_u8, threads=8):
        target = np.ascontiguousarray(target_u8, np.uint8)
        assert target.shape == (params.height, params.width, 3)
        loss = C.c_double(0)
        grad = np.zeros((self.nT, 3), np.float64)
        rows = n_rows(params)
        img = np.zeros((rows, params.width, 3), np.float32)
        lib().orc_render_grad(C.c_void_p(self.h), C.byref(params), _ptr(target), C.byref(loss), _ptr(grad), _ptr(img), threads)
        return loss.value, grad, img

    def object_prims(self):
        boxes = np.zeros((self.nO, 6), np.float32)
        pos = np.zeros((self.nO, 3), np.float32)
        lib().orc_object_prims(C.c_void_p(self.h), _ptr(boxes), _ptr(pos))
        return boxes, pos


def bvh_build(boxes, pos, leaf_size=4):
    boxes = np.ascontiguousarray(boxes, np.float32).reshape(-1, 6)
    pos = np.ascontiguousarray(pos, np.float32).reshape(-1, 3)
    n = boxes.shape[0]
    max_nodes = max(1, 2 * n)
    nb = np.zeros((max_nodes, 6), np.float32)
    ni = np.zeros((max_nodes, 3), np.int32)
    perm = np.zeros(n, np.int32)
    k = lib().orc_bvh_build(_ptr(boxes), _ptr(pos), n, leaf_size, _ptr(nb), _ptr(ni), max_nodes, _ptr(perm))
    return nb[:k], ni[:k], perm


def philox(ctr, key):
    c = np.array(ctr, np.uint32)
    k = np.array(key, np.uint32)
    o = np.zeros(4, np.uint32)
    lib().orc_philox(_ptr(c), _ptr(k), _ptr(o))
    return o


def uniforms(seed, pixel, sample, bounce, stream):
    u = np.zeros(6, np.float32)
    lib().orc_uniforms(C.c_uint32(seed & 0xFFFFFFFF), C.c_uint32(seed >> 32), C.c_uint32(pixel), C.c_uint32(sample),
                       C.c_uint32(bounce), C.c_uint32(stream), _ptr(u))
    return u


def sincos2pi(u):
    s, c = C.c_float(0), C.c_float(0)
    lib().orc_sincos2pi(C.c_float(u), C.byref(s), C.byref(c))
    return s.value, c.value


def hardware_threads():
    return lib().orc_hardware_threads()


def ref_loader_dump(obj_path, mtl_path):
    """Run the REFERENCE's own loader (oracle/_ref, built from /root/reference) and parse its table dump."""
    out = subprocess.check_output([REF_DUMP, obj_path, mtl_path]).decode().splitlines()
    it = iter(out)
    res = {}

    def fl(x):
        return np.float32(float.fromhex(x))

    hdr = next(it).split()
    res["V"] = np.array([[fl(x) for x in next(it).split()] for _ in range(int(hdr[1]))], np.float32).reshape(-1, 3)
    hdr = next(it).split()
    res["N"] = np.array([[fl(x) for x in next(it).split()] for _ in range(int(hdr[1]))], np.float32).reshape(-1, 3)
    hdr = next(it).split()
    mats = []
    for _ in range(int(hdr[1])):
        name, vals = next(it).split(" | ")
        v = vals.split()
        mats.append({"name": name, "kd": [fl(x) for x in v[0:3]], "ks": [fl(x) for x in v[3:6]], "ke": [fl(x) for x in v[6:9]],
                     "ns": fl(v[9]), "ni": fl(v[10]), "d": fl(v[11]), "illum": int(v[12])})
    res["M"] = mats
    hdr = next(it).split()
    res["F"] = np.array([[int(x) for x in next(it).split()] for _ in range(int(hdr[1]))], np.int32).reshape(-1, 4)
    return res

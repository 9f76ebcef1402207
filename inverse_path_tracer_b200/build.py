"""In-tree build of libipt_b200.so (sm_100a only) and the drop-in library names.

    python -m inverse_path_tracer_b200.build

Produces inverse_path_tracer_b200/lib/libipt_b200.so and the two names the reference's ipt_cuda.py
loads (ipt_cuda.py:7-8): build/libpt.so and build/libipt.so, as symlinks to the one library.

Flags that matter for correctness: -fmad=false (device) and -ffp-contract=off (host) -- the canon of
csrc/ipt_math.cuh fuses only where it says fmaf().
"""
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIBDIR = os.path.join(PKG, "lib")
OBJDIR = os.path.join(LIBDIR, "obj")
LIB = os.path.join(LIBDIR, "libipt_b200.so")

SOURCES = ["api.cu", "inverse.cu", "paths.cu", "trace.cu", "wavefront.cu", "bvh_build.cu", "scene_host.cu", "png_io.cpp"]
HEADERS = ["ipt_math.cuh", "intersect.cuh", "shading.cuh", "device_scene.h", "kernels.h", "scene_host.h", "scene_handle.h", "nvtx_ranges.h", os.path.join(ROOT, "include", "ipt_b200.h")]

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
              "-Xcompiler", "-fPIC,-ffp-contract=off,-O2,-Wall,-Wno-unused-function", "-Xptxas", "-v"]


def nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; libipt_b200.so cannot be built")
    return exe


def _newer(a, deps):
    if not os.path.exists(a):
        return True
    t = os.path.getmtime(a)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    os.makedirs(OBJDIR, exist_ok=True)
    hdrs = [h if os.path.isabs(h) else os.path.join(CSRC, h) for h in HEADERS]
    objs = []
    logs = []
    for src in SOURCES:
        sp = os.path.join(CSRC, src)
        obj = os.path.join(OBJDIR, os.path.splitext(src)[0] + ".o")
        objs.append(obj)
        if force or _newer(obj, [sp] + hdrs + [os.path.abspath(__file__)]):
            cmd = [nvcc()] + NVCC_FLAGS + ["-c", sp, "-o", obj]
            r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
            logs.append(r.stdout)
            if r.returncode != 0:
                sys.stderr.write(r.stdout)
                raise RuntimeError("nvcc failed on " + src)
            if verbose:
                sys.stderr.write(r.stdout)
    if force or _newer(LIB, objs):
        cmd = [nvcc(), "-shared", "-o", LIB] + objs + ["-lz", "-gencode", "arch=compute_100a,code=sm_100a"]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout)
            raise RuntimeError("link failed")
    if logs:
        with open(os.path.join(LIBDIR, "ptxas.log"), "w") as f:
            f.write("\n".join(logs))
    # the names the reference's ipt_cuda.py dlopens, cwd-relative to the repo root
    bdir = os.path.join(ROOT, "build")
    os.makedirs(bdir, exist_ok=True)
    for name in ("libpt.so", "libipt.so"):
        link = os.path.join(bdir, name)
        if os.path.islink(link) or os.path.exists(link):
            os.remove(link)
        os.symlink(os.path.relpath(LIB, bdir), link)
    return LIB


def build_variant(name, defines, sources=("paths.cu",)):
    """Developer A/B build: lib/ab/<name>.so = the library with `sources` recompiled under extra -D flags
    (select it with IPT_B200_LIB=...; tools/ab_variants.py).  The shipped library is not touched."""
    build()
    abdir = os.path.join(LIBDIR, "ab")
    os.makedirs(abdir, exist_ok=True)
    objs = []
    for src in SOURCES:
        base = os.path.splitext(src)[0]
        if src in sources:
            obj = os.path.join(OBJDIR, "ab_%s_%s.o" % (name, base))
            cmd = [nvcc()] + NVCC_FLAGS + ["-D" + d for d in defines] + ["-c", os.path.join(CSRC, src), "-o", obj]
            r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
            if r.returncode != 0:
                sys.stderr.write(r.stdout)
                raise RuntimeError("nvcc failed on %s (%s)" % (src, name))
            with open(os.path.join(abdir, name + ".ptxas.log"), "w") as f:
                f.write(r.stdout)
        else:
            obj = os.path.join(OBJDIR, base + ".o")
        objs.append(obj)
    out = os.path.join(abdir, name + ".so")
    r = subprocess.run([nvcc(), "-shared", "-o", out] + objs + ["-lz", "-gencode", "arch=compute_100a,code=sm_100a"],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("link failed (%s)" % name)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))

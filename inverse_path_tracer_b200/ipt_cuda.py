"""Host-side mirror of the reference's Python entry points (ipt_cuda.py), same names and argument
meaning, running on libipt_b200.so.

The reference module itself also runs UNMODIFIED against this library: `python -m
inverse_path_tracer_b200.build` creates ./build/libpt.so and ./build/libipt.so (the names
ipt_cuda.py:7-8 loads).  This mirror exists so the path can be used without the reference checkout
and without its cwd-relative assumptions (ASSET_ROOT), and it adds the runtime parameters.

reference -> here
  load_params(filename)                 ipt_cuda.py:91-107   same tuple for loadScene
  load_scene(scenefile)                 ipt_cuda.py:109-113  -> (n_t, scene_ptr)
  generate_files(n)                     ipt_cuda.py:115-134  writes scenes/i.txt + imgs/i.png
  generate_data(scenefile, imgfile)     ipt_cuda.py:136-165  -> (w, pixel, light, labels)
  render_with_materials(scenefile, imgfile, materials)  ipt_cuda.py:167-183
"""
import ctypes as C
import os

import numpy as np

from . import native

(Cube, Sphere, Cornell, Other) = (0, 1, 2, 3)
ASSET_ROOT = ""   # prepended to relative OBJ/MTL paths ("" = cwd-relative like the reference)

c_float_p = C.POINTER(C.c_float)


def rand_mtl():
    return "*Kd %s %s %s*" % (np.random.uniform(), np.random.uniform(), np.random.uniform())


def to_string(shp=None, pos=None, ori=None, scl=None, obj_file=None, mtl_file=None):
    s = ""
    if pos is not None:
        s += "POS %s %s %s\n" % (pos[0], pos[1], pos[2])
    if ori is not None:
        s += "ORI %s %s %s\n" % (ori[0], ori[1], ori[2])
    if scl is not None:
        s += "SCL %s %s %s\n" % (scl[0], scl[1], scl[2])
    if shp is Cube:
        obj_file = "./shapes/cube.obj"
        mtl_file = rand_mtl() if mtl_file is None else mtl_file
    elif shp is Sphere:
        obj_file = "./shapes/sphere.obj"
        mtl_file = rand_mtl() if mtl_file is None else mtl_file
    elif shp is Cornell:
        obj_file = "./CornellBox/CornellBox-Empty-CO.obj"
        mtl_file = "./CornellBox/CornellBox-Empty-CO.mtl"
    assert obj_file is not None and mtl_file is not None
    s += "OBJ %s\n" % obj_file
    s += "MTL %s\n" % mtl_file
    return s


def from_string(string):
    """One OBJECT block -> dict(pos, ori, scl, obj, mtl); defaults 0 / 0 / 1 (ipt_cuda.py:39-59)."""
    o = {"pos": None, "ori": None, "scl": None, "obj": None, "mtl": None}
    for line in string.split("\n"):
        items = line.strip().split(" ")
        tok, vals = items[0], items[1:]
        if tok == "POS":
            o["pos"] = [float(x) for x in vals]
        elif tok == "ORI":
            o["ori"] = [float(x) for x in vals]
        elif tok == "SCL":
            o["scl"] = [float(x) for x in vals]
        elif tok == "OBJ":
            o["obj"] = vals[0]
        elif tok == "MTL":
            o["mtl"] = " ".join(vals)
    if o["pos"] is None:
        o["pos"] = [0.0] * 3
    if o["ori"] is None:
        o["ori"] = [0.0] * 3
    if o["scl"] is None:
        o["scl"] = [1.0] * 3
    assert o["obj"] is not None and o["mtl"] is not None
    return o


def parse_scene_file(filename):
    """scene file -> list of object dicts (ipt_cuda.py:91-107)."""
    objs, cur = [], ""
    with open(filename, "r") as f:
        for line in f.readlines():
            line = line.strip()
            if line == "OBJECT":
                if len(cur) > 0:
                    objs.append(from_string(cur))
                cur = ""
            else:
                cur += line + "\n"
    objs.append(from_string(cur))
    return objs


def dereference(py_objects):
    n = len(py_objects)
    keep = [[(C.c_float * 3)(*o[k]) for o in py_objects] for k in ("pos", "ori", "scl")]
    poss, oris, scls = [(c_float_p * n)(*[C.cast(a, c_float_p) for a in col]) for col in keep]
    objs = (C.c_char_p * n)(*[o["obj"].encode("utf-8") for o in py_objects])
    mtls = (C.c_char_p * n)(*[o["mtl"].encode("utf-8") for o in py_objects])
    params = (poss, oris, scls, objs, mtls, n)
    params[0]._keep = keep   # keep the float[3] blocks alive as long as the pointer array
    return params


def load_params(filename):
    return dereference(parse_scene_file(filename))


def load_scene(scenefile):
    L = native.lib()
    L.iptSetAssetRoot(ASSET_ROOT.encode())
    params = load_params(scenefile)
    scene_ptr = C.c_void_p(0)
    n_t = L.loadScene(*params, C.pointer(scene_ptr))
    if not scene_ptr.value:
        raise RuntimeError("loadScene failed: " + native.last_error())
    return n_t, scene_ptr


def generate_files(n, scene_dir="scenes", img_dir="imgs"):
    L = native.lib()
    os.makedirs(scene_dir, exist_ok=True)
    os.makedirs(img_dir, exist_ok=True)
    for i in range(n):
        with open("%s/%d.txt" % (scene_dir, i), "w") as f:
            f.write("OBJECT\n")
            f.write(to_string(shp=Cornell, pos=np.array([0, 0, 4]), scl=np.ones(3) * 2))
            f.write("OBJECT\n")
            f.write(to_string(shp=Cube, pos=np.array([0, -1.5, 4])))
        n_t, scene_ptr = load_scene("%s/%d.txt" % (scene_dir, i))
        L.createImage(scene_ptr, ("%s/%d.png" % (img_dir, i)).encode("utf-8"))
        L.freeScene(scene_ptr)


def generate_data(scenefile, imgfile):
    L = native.lib()
    len_data = 7
    n_t, scene_ptr = load_scene(scenefile)
    size = (n_t + 1) * n_t
    data = (C.c_float * (size * len_data))(0)
    L.createGraph(scene_ptr, imgfile.encode("utf-8"), data)
    labels = (C.c_float * (n_t * 3))(0)
    L.getMaterials(scene_ptr, labels)
    labels = np.ctypeslib.as_array(labels).astype(np.float64).reshape(n_t, 3)
    L.freeScene(scene_ptr)
    data = np.ctypeslib.as_array(data).astype(np.float64)
    w = data[:size].reshape(n_t + 1, n_t)
    pixel = data[size:size * 4].reshape(n_t + 1, n_t, 3)
    assert not np.isnan(pixel).any()
    light = data[size * 4:].reshape(n_t + 1, n_t, 3)
    return w, pixel, light, labels


def render_with_materials(scenefile, imgfile, materials):
    L = native.lib()
    n_t, scene_ptr = load_scene(scenefile)
    m = materials.numpy() if hasattr(materials, "numpy") else np.asarray(materials)
    m = np.ascontiguousarray(m, np.float32)
    L.setMaterials(scene_ptr, m.ctypes.data_as(C.c_void_p))
    L.createImage(scene_ptr, imgfile.encode("utf-8"))
    L.freeScene(scene_ptr)

"""CPU tests of the product's host side: loader, BVH builder, PNG codec, C-ABI surface, error
behaviour.  No compute call is made (there is no CPU fallback); scenes are loaded host-only."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ASSETS, GOLDEN, ROOT, scene_objects


def _host_scene(ipt, objects, **kw):
    return ipt.Scene(objects, asset_root=ASSETS, device=-2, **kw)


def test_library_exports_every_declared_symbol(ipt):
    from inverse_path_tracer_b200 import native
    hdr = open(os.path.join(ROOT, "include", "ipt_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{]*\)\s*;", hdr)
    assert {"loadScene", "freeScene", "createImage", "createGraph", "getMaterials", "setMaterials"} <= set(declared)
    L = native.lib()
    for name in declared:
        assert hasattr(L, name), "libipt_b200.so does not export %s" % name
    assert set(declared) == set(native.EXPORTS)


def test_drop_in_library_names_resolve_to_one_library(ipt):
    from inverse_path_tracer_b200 import native
    for name in ("libpt.so", "libipt.so"):       # ipt_cuda.py:7-8
        p = os.path.join(ROOT, "build", name)
        if not os.path.exists(p):
            pytest.skip("build/ symlinks are created by `python -m inverse_path_tracer_b200.build`")
        assert os.path.realpath(p) == os.path.realpath(native.library_path())
        L = C.CDLL(p)
        for sym in ("loadScene", "freeScene", "createImage", "createGraph", "getMaterials", "setMaterials"):
            assert hasattr(L, sym)


@pytest.mark.parametrize("name", ["scene0", "cornell", "cornell_sphere", "rotated_cubes"])
def test_loader_tables_bit_identical_to_oracle(ipt, oracle, name):
    objects = oracle.parse_scene_file(os.path.join(ASSETS, "scenes", "0.txt")) if name == "scene0" else scene_objects(name)
    g, o = _host_scene(ipt, objects), oracle.OracleScene(objects, ASSETS)
    assert (g.nT, g.nE, g.nO) == (o.nT, o.nE, o.nO)
    a, b = g.triangle_table(), o.triangle_table()
    for k in a:
        assert a[k].tobytes() == b[k].tobytes(), k
    assert np.array_equal(g.camera(), o.camera())
    assert np.array_equal(g.get_materials(), o.get_materials())


def _fuzz_cases():
    path = os.path.join(GOLDEN, "loader_fuzz.npz")
    return [str(c) for c in np.load(path)["cases"]] if os.path.exists(path) else []


@pytest.mark.parametrize("case", _fuzz_cases())
def test_loader_fuzz_product_against_reference_tables(ipt, oracle, case):
    """the product's own OBJ/MTL parser on the fuzz files (tools/make_loader_fuzz.py): equal to the tables the
    reference's loader produced, and bit-identical to the oracle's table in every field"""
    from test_oracle import check_fuzz_tables, fuzz_objects
    z = np.load(os.path.join(GOLDEN, "loader_fuzz.npz"))
    g = _host_scene(ipt, fuzz_objects(case))
    check_fuzz_tables(g.triangle_table(), g.nT, z[case + "_V"], z[case + "_N"], z[case + "_F"], z[case + "_M"])
    o = oracle.OracleScene(fuzz_objects(case), ASSETS)
    a, b = g.triangle_table(), o.triangle_table()
    assert (g.nT, g.nE) == (o.nT, o.nE)
    for k in a:
        assert a[k].tobytes() == b[k].tobytes(), k


def test_set_get_materials_roundtrip_and_handle_validation(ipt):
    from inverse_path_tracer_b200 import native
    g = _host_scene(ipt, scene_objects("cornell"))
    m = np.random.RandomState(0).uniform(0, 1, (g.nT, 3)).astype(np.float32)
    g.set_materials(m)
    assert np.array_equal(g.get_materials(), m)
    L = native.lib()
    bogus = C.c_void_p(0x1234)
    out = np.full(3, -1, np.float32)
    L.getMaterials(bogus, out.ctypes.data_as(C.c_void_p))     # must not crash, must not write
    assert (out == -1).all()
    L.freeScene(bogus)
    L.freeScene(g.h)
    L.freeScene(g.h)                                          # double free is ignored
    g.h = C.c_void_p(0)


def test_gpu_builder_is_required_for_large_inputs_without_a_device(ipt):
    import torch
    from inverse_path_tracer_b200 import native
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    pos = np.random.RandomState(0).uniform(-1, 1, (3000, 3)).astype(np.float32)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        native.build_bvh(np.concatenate([pos, pos], 1), pos)          # auto: >= 2048 primitives -> CUDA kernels


def test_no_cpu_fallback(ipt):
    g = _host_scene(ipt, scene_objects("cornell"))
    p = ipt.default_params(width=8, height=8, spp=1)
    with pytest.raises(RuntimeError, match="host-only"):
        g.render(p)
    with pytest.raises(RuntimeError, match="host-only"):
        g.trace(np.zeros((1, 6), np.float32))
    with pytest.raises(RuntimeError, match="host-only"):
        g.render_grad(p, np.zeros((8, 8, 3), np.uint8))
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback|CUDA"):
            ipt.Scene(scene_objects("cornell"), asset_root=ASSETS)


def test_load_errors_are_reported_not_fatal(ipt):
    from inverse_path_tracer_b200 import native
    bad = [{"pos": [0, 0, 0], "ori": [0, 0, 0], "scl": [1, 1, 1], "obj": "./does/not/exist.obj", "mtl": "*Kd 1 1 1*"}]
    with pytest.raises(RuntimeError, match="Object File was not able to be opened"):   # reference: printf + exit(1)
        _host_scene(ipt, bad)
    assert "exist.obj" in native.last_error()


def test_default_params_are_the_reference_macros(ipt):
    p = ipt.default_params()
    assert (p.width, p.height, p.spp, p.max_depth) == (500, 500, 100, 0)     # scene.h:8-10, no depth limit
    assert abs(p.p_rr - 0.9) < 1e-7                                           # scene.h:11


@pytest.mark.parametrize("n,seed", [(0, 0), (1, 0), (3, 1), (4, 2), (5, 3), (8, 4), (64, 5), (1000, 6), (20000, 7)])
def test_bvh_builder_layout_bit_exact_vs_oracle(ipt, oracle, n, seed):
    from inverse_path_tracer_b200 import native
    rng = np.random.RandomState(seed)
    pos = rng.uniform(-3, 3, (n, 3)).astype(np.float32)
    ext = rng.uniform(0, 0.3, (n, 3)).astype(np.float32)
    boxes = np.concatenate([pos - ext, pos + ext], 1).astype(np.float32)
    nb_g, ni_g, perm_g = native.build_bvh(boxes, pos, builder=1)   # host execution (CPU suite); the CUDA kernels: test_gpu_parity.py
    nb_o, ni_o, perm_o = oracle.bvh_build(boxes, pos)
    assert np.array_equal(perm_g, perm_o) and np.array_equal(ni_g, ni_o) and nb_g.tobytes() == nb_o.tobytes()


def test_bvh_builder_degenerate_inputs(ipt, oracle):
    from inverse_path_tracer_b200 import native
    rng = np.random.RandomState(9)
    cases = []
    same = np.tile(np.array([[1.0, 2.0, 3.0]], np.float32), (37, 1))                      # all centroids equal -> median fallback
    cases.append(same)
    grid = np.stack(np.meshgrid(np.arange(6), np.arange(5), np.arange(4), indexing="ij"), -1).reshape(-1, 3).astype(np.float32)
    cases.append(grid)                                                                    # many ties on the split plane
    cases.append(rng.permutation(grid))
    line = np.zeros((50, 3), np.float32)
    line[:, 1] = np.repeat(np.arange(10), 5)                                              # zero extent on two axes
    cases.append(line)
    dup = rng.uniform(-1, 1, (40, 3)).astype(np.float32)
    cases.append(np.concatenate([dup, dup, dup[:7]]))                                     # duplicates
    for pos in cases:
        boxes = np.concatenate([pos - 0.01, pos + 0.02], 1).astype(np.float32)
        a, b = native.build_bvh(boxes, pos, builder=1), oracle.bvh_build(boxes, pos)
        assert np.array_equal(a[2], b[2]) and np.array_equal(a[1], b[1]) and a[0].tobytes() == b[0].tobytes()
    # signed zeros: bounds are reduced under the total order -0 < +0
    pos = np.array([[0.0, -0.0, 0.0], [-0.0, 0.0, 1.0], [0.0, 0.0, -0.0], [-0.0, -0.0, 2.0], [1.0, 0.0, 0.0], [-0.0, 1.0, -0.0]], np.float32)
    boxes = np.concatenate([pos, pos], 1)
    a, b = native.build_bvh(boxes, pos, builder=1), oracle.bvh_build(boxes, pos)
    assert a[0].tobytes() == b[0].tobytes() and np.array_equal(a[2], b[2])


def test_bvh_builder_on_scene_triangles_and_objects(ipt, oracle):
    from inverse_path_tracer_b200 import native
    o = oracle.OracleScene(scene_objects("cornell_sphere"), ASSETS)
    t = o.triangle_table()
    boxes = np.concatenate([t["verts"].min(1), t["verts"].max(1)], 1)
    a, b = native.build_bvh(boxes, t["center"], builder=1), oracle.bvh_build(boxes, t["center"])
    assert np.array_equal(a[2], b[2]) and np.array_equal(a[1], b[1]) and a[0].tobytes() == b[0].tobytes()
    ob, op = oracle.OracleScene(scene_objects("rotated_cubes"), ASSETS).object_prims()     # 5 objects > leafSize: a real split
    a, b = native.build_bvh(ob, op), oracle.bvh_build(ob, op)
    assert a[1].shape[0] == 3 and np.array_equal(a[2], b[2]) and np.array_equal(a[1], b[1]) and a[0].tobytes() == b[0].tobytes()


def test_png_roundtrip(ipt, tmp_path):
    from PIL import Image
    from inverse_path_tracer_b200 import native
    L = native.lib()
    rng = np.random.RandomState(0)
    img = rng.randint(0, 256, (37, 53, 3)).astype(np.uint8)
    p = str(tmp_path / "a.png")
    assert L.iptWritePNG(p.encode(), img.ctypes.data_as(C.c_void_p), 53, 37) == 0
    assert np.array_equal(np.asarray(Image.open(p)), img)
    for mode, arr in (("RGB", img), ("RGBA", np.dstack([img, rng.randint(0, 256, (37, 53, 1)).astype(np.uint8)])), ("L", img[:, :, 0]),
                      ("P", None)):
        q = str(tmp_path / ("b_%s.png" % mode))
        if mode == "P":
            Image.fromarray(img).convert("P", palette=Image.ADAPTIVE, colors=32).save(q)
            want = np.asarray(Image.open(q).convert("RGB"))
        else:
            Image.fromarray(arr, mode).save(q)
            want = np.asarray(Image.open(q).convert("RGB"))
        out = np.zeros((37, 53, 3), np.uint8)
        w, h = C.c_int(0), C.c_int(0)
        assert L.iptReadPNG(q.encode(), out.ctypes.data_as(C.c_void_p), out.nbytes, C.byref(w), C.byref(h)) == 0, native.last_error()
        assert (w.value, h.value) == (53, 37) and np.array_equal(out, want)
    assert L.iptReadPNG(b"/nonexistent.png", None, 0, None, None) == -1 and "cannot open" in native.last_error()


def test_scene_file_mirror_matches_reference_grammar(ipt, oracle, tmp_path):
    from inverse_path_tracer_b200 import ipt_cuda
    for i in (0, 42):
        f = os.path.join(ASSETS, "scenes", "%d.txt" % i)
        assert ipt_cuda.parse_scene_file(f) == oracle.parse_scene_file(f)
    # to_string/from_string round trip with defaults 0 / 0 / 1 (ipt_cuda.py:17-59)
    s = ipt_cuda.to_string(shp=ipt_cuda.Cube, pos=[1, 2, 3], mtl_file="*Kd 0.1 0.2 0.3*")
    o = ipt_cuda.from_string(s)
    assert o["pos"] == [1.0, 2.0, 3.0] and o["ori"] == [0.0] * 3 and o["scl"] == [1.0] * 3
    assert o["obj"] == "./shapes/cube.obj" and o["mtl"] == "*Kd 0.1 0.2 0.3*"
    poss, oris, scls, objs, mtls, n = ipt_cuda.load_params(os.path.join(ASSETS, "scenes", "0.txt"))
    assert n == 2 and [poss[1][k] for k in range(3)] == [0.0, -1.5, 4.0] and objs[0] == b"./CornellBox/CornellBox-Empty-CO.obj"


def test_graph_compress_matches_oracle_compress(ipt, oracle):
    """DataWrapper::compress (inv_scene.h:87-115) on synthetic raw sums, host-only"""
    from inverse_path_tracer_b200 import native
    sc = oracle.OracleScene.from_file(os.path.join(ASSETS, "scenes", "0.txt"), ASSETS)
    target = np.random.RandomState(1).randint(0, 256, (24, 24, 3)).astype(np.uint8)
    raw, data, _ = sc.create_graph(oracle.make_params(24, 24, 4, seed=3), target, threads=4)
    assert native.graph_compress(raw, sc.nT).tobytes() == data.tobytes()


def test_gcn_consumer_matches_edge_list_message_passing():
    """inverse_path_tracer_b200/gcn.py restates the reference's DGL network (ipt.py:28-84) with a dense weight matrix:
    the dense message pass must equal the explicit per-edge sum, graph construction must follow ipt.py:67-82
    (threshold, row normalisation, eye row -> node inputs), and a short fit must reduce the loss."""
    import torch
    from inverse_path_tracer_b200 import gcn
    rng = np.random.RandomState(0)
    nT = 12
    w = rng.uniform(0, 1, (nT + 1, nT)) * (rng.uniform(0, 1, (nT + 1, nT)) < 0.4)
    w[3] = 5e-4                       # a row entirely under P_MIN stays zero after thresholding
    pixel = rng.uniform(0, 1, (nT + 1, nT, 3))
    labels = rng.uniform(0, 1, (nT, 3))
    W, x = gcn.build_graph(w, pixel, None)
    assert W.shape == (nT, nT) and x.shape == (nT, 3)
    assert torch.equal(x, torch.tensor(pixel[-1], dtype=torch.float32))
    assert float(W[3].abs().sum()) == 0.0
    rows = W.sum(-1)
    assert all(abs(float(r) - 1.0) < 1e-6 or float(r) == 0.0 for r in rows)
    Wt, xt = gcn.build_graph(torch.tensor(w, dtype=torch.float32), torch.tensor(pixel, dtype=torch.float32), None)   # tensor inputs
    assert torch.allclose(Wt, W, atol=1e-6) and torch.equal(xt, x)
    torch.manual_seed(0)
    layer = gcn.MPL(5, 7, None)
    h = torch.randn(nT, 5)
    dst, src = np.nonzero(W.numpy())
    reduced = torch.zeros(nT, 5)
    for d, s_ in zip(dst, src):       # DGL: msg = h[src] * w(edge), summed at dst; self loops carry weight 0
        reduced[d] += h[s_] * W[d, s_]
    want = layer.linear(torch.cat((h, reduced), -1))
    assert torch.allclose(layer(W, h), want, atol=1e-6)
    model = gcn.GCN(3, 3)
    out1 = model(W, x)
    outb = model(torch.stack([W, W]), torch.stack([x, x]))
    assert out1.shape == (nT, 3) and torch.allclose(outb[0], out1, atol=1e-6) and float(out1.min()) > 0 and float(out1.max()) < 1
    data = [(w, pixel, None, labels)]
    _, l_short = gcn.fit(data, epochs=1, lr=1e-2, device="cpu")
    _, l_long = gcn.fit(data, epochs=150, lr=1e-2, device="cpu")
    assert l_long < 0.7 * l_short

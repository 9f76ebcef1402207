"""GPU parity tests proper: the CUDA path (through the C ABI) against the CPU oracle on the same
seeded inputs.  Bars (BASELINE.json north_star): hit ids bit-exact, t within 1e-5 relative (we
require bit-exact), images within RMSE tolerance (we require bit-exact for diffuse scenes because
both sides implement the same FP32 canon), edge statistics / gradients within 1e-3 relative."""
import json
import os

import numpy as np
import pytest

from conftest import ASSETS, GOLDEN, random_rays, scene_objects

pytestmark = pytest.mark.gpu


def _pair(ipt, oracle, objects, **kw):
    return ipt.Scene(objects, asset_root=ASSETS, **kw), oracle.OracleScene(objects, ASSETS)


def _scene0(oracle):
    return oracle.parse_scene_file(os.path.join(ASSETS, "scenes", "0.txt"))


def adversarial_rays(osc):
    """edge-grazing, shared-diagonal, coplanar ceiling/light seam (y = 2) and self-hit (t ~ EPSILON) rays"""
    tab = osc.triangle_table()
    v = tab["verts"].astype(np.float64)
    rng = np.random.RandomState(11)
    rays = []
    eye = np.array([0.0, 0.0, 0.0])
    for i in range(v.shape[0]):
        for j in range(3):
            a, b = v[i, j], v[i, (j + 1) % 3]
            for s in (0.0, 0.25, 0.5, 1.0):   # points exactly on edges and vertices
                q = a + s * (b - a)
                for o in (eye, np.array([0.3, -0.4, 2.5]), rng.uniform(-1.5, 1.5, 3) + np.array([0, 0, 4.0])):
                    d = q - o
                    n = np.linalg.norm(d)
                    if n > 1e-6:
                        rays.append(np.concatenate([o, d / n]))
        c = v[i].mean(0)
        nrm = tab["normal"][i].astype(np.float64)
        for eps in (0.0, 5e-3, 9.9e-3, 1.0e-2, 1.01e-2, 2e-2):   # origins on / just off the surface
            for k in range(3):
                d = rng.normal(size=3)
                d /= np.linalg.norm(d)
                rays.append(np.concatenate([c + eps * nrm, d]))
                rays.append(np.concatenate([c - eps * nrm, d]))
        # grazing directions (|n.d| around MIN_DOT = 1e-4)
        t1 = v[i, 1] - v[i, 0]
        t1 /= np.linalg.norm(t1)
        for s in (0.0, 5e-5, 1e-4, 1.5e-4, -1e-4):
            d = t1 + s * nrm
            d /= np.linalg.norm(d)
            rays.append(np.concatenate([c - 0.5 * t1 + 0.3 * nrm * (s != 0), d]))
    return np.array(rays, np.float32)


@pytest.mark.parametrize("scene,mode", [("scene0", 0), ("scene0", 2), ("cornell", 0), ("rotated_cubes", 0), ("rotated_cubes", 2),
                                        ("cornell_sphere", 0), ("cornell_sphere", 1)])
def test_closest_hit_bit_exact(ipt, oracle, scene, mode):
    objects = _scene0(oracle) if scene == "scene0" else scene_objects(scene)
    g, o = _pair(ipt, oracle, objects, traversal_mode=mode)
    assert g.nT == o.nT and g.nE == o.nE
    assert bool(g.uses_bvh) == (mode == 2 or (mode == 0 and g.nT > 128))
    n = 2_000_000 if g.nT <= 64 else 200_000
    rays = np.concatenate([random_rays(n, 5), adversarial_rays(o)])
    tri_g, t_g = g.trace(rays)
    tri_o, t_o = o.trace(rays)
    assert np.array_equal(tri_g, tri_o), "hit ids differ at %s" % np.nonzero(tri_g != tri_o)[0][:10]
    assert t_g.tobytes() == t_o.tobytes()
    assert (tri_o >= 0).mean() > 0.5


def slab_adversarial_rays(osc, seed=3):
    """Rays against the traversal's box test (quantised records: plane distance = fma(2^23 + q, scale/d, (off - p)/d - 2^23 scale/d),
    intersect.cuh): origins 10 .. 1e4 units away aimed at points on triangles (cancellation in (off - p)/d), directions exactly
    parallel to an axis or with one / two components of 1e-4 .. 1e-30 (clamped reciprocals, huge plane distances), origins exactly on
    triangle vertices and inside box face planes.

    Known limit (found with this generator at 1e5 units, session r02_s28): the FP32 triangle test itself accepts a
    hit whose t is off by ~2^-24 (|n.p| + |k|) / |n.d|; from 1e5 units away that is 0.008 along the ray, enough to accept a triangle the
    ray misses by 3e-4 -- outside its padded box (pad 1e-5 (1 + |coordinate|)), so the traversal returns the next hit instead.  One of
    39600 rays, identically for float, quantised and 4-wide node records: a property of fixed padding, not of the record format."""
    tab = osc.triangle_table()
    v = tab["verts"].astype(np.float64)
    rng = np.random.RandomState(seed)
    nT = v.shape[0]
    rays = []
    pick = rng.randint(0, nT, 600)
    for i in pick:
        w = rng.dirichlet([1, 1, 1])
        q = w @ v[i]
        for dist in (10.0, 1e2, 1e3, 1e4):
            u = rng.normal(size=3)
            u /= np.linalg.norm(u)
            rays.append(np.concatenate([q - dist * u, u]))
            k = rng.randint(3)                                   # approach along an axis, exactly and almost
            e = np.zeros(3); e[k] = rng.choice([-1.0, 1.0])
            rays.append(np.concatenate([q - dist * e, e]))
            for tiny in (1e-4, 1e-7, 1e-12, 1e-30):
                d = e.copy(); d[(k + 1) % 3] = tiny * rng.choice([-1, 1]); d[(k + 2) % 3] = tiny * rng.choice([-1, 0, 1])
                d /= np.linalg.norm(d)
                rays.append(np.concatenate([q - min(dist, 1e3) * d, d]))
        for j in range(3):                                       # origins exactly on vertices, axis-parallel directions
            for k in range(3):
                for sgn in (-1.0, 1.0):
                    e = np.zeros(3); e[k] = sgn
                    rays.append(np.concatenate([v[i, j], e]))
                    rays.append(np.concatenate([v[i, j] - 3.0 * e, e]))    # ... and running through the vertex from outside
    return np.array(rays, np.float32)


@pytest.mark.parametrize("scene,mode", [("cornell_sphere", 0), ("cornell_sphere", 2), ("rotated_cubes", 2)])
def test_traversal_box_test_adversarial_rays(ipt, oracle, scene, mode):
    """the BVH path must return the brute-force closest hit bit for bit whatever the box test does with far origins, axis-parallel
    directions and origins on box faces (a conservative test may only visit more nodes)"""
    objects = scene_objects(scene)
    g, o = _pair(ipt, oracle, objects, traversal_mode=mode)
    assert g.uses_bvh
    rays = slab_adversarial_rays(o)
    tri_g, t_g = g.trace(rays)
    tri_o, t_o = o.trace(rays)
    bad = np.nonzero((tri_g != tri_o) | (t_g.view(np.uint32) != t_o.view(np.uint32)))[0]
    assert bad.size == 0, "first mismatches: %s" % [(int(b), rays[b].tolist(), int(tri_g[b]), int(tri_o[b]), float(t_g[b]), float(t_o[b])) for b in bad[:3]]
    assert (tri_o >= 0).mean() > 0.5


def test_trace_edge_cases(ipt, oracle):
    g, o = _pair(ipt, oracle, scene_objects("cornell"))
    tri, t = g.trace(np.zeros((0, 6), np.float32))
    assert tri.shape == (0,)
    rays = np.array([[0, 0, 0, 0, 0, -1], [0, 0, 0, 0, 0, 0], [0, 0, 4, 0, 1, 0], [np.nan, 0, 0, 0, 0, 1]], np.float32)
    tri_g, t_g = g.trace(rays)
    tri_o, t_o = o.trace(rays)
    assert np.array_equal(tri_g, tri_o) and t_g.tobytes() == t_o.tobytes()
    assert tri_g[0] == -1 and np.isinf(t_g[0])


@pytest.mark.parametrize("scene,w,h,spp,depth,gs", [("scene0", 64, 48, 16, 0, 64), ("scene0", 40, 40, 70, 5, 32), ("cornell", 256, 256, 16, 5, 64),
                                                    ("rotated_cubes", 48, 48, 8, 8, 3)])
def test_forward_image_bit_exact(ipt, oracle, scene, w, h, spp, depth, gs):
    objects = _scene0(oracle) if scene == "scene0" else scene_objects(scene)
    g, o = _pair(ipt, oracle, objects)
    pg = ipt.default_params(width=w, height=h, spp=spp, max_depth=depth, seed=1234, group_samples=gs)
    po = oracle.make_params(w, h, spp, max_depth=depth, seed=1234, group_samples=gs)
    img_g, img8_g, cnt_g = g.render(pg, want_u8=True)
    img_o, img8_o, cnt_o = o.render(po, want_u8=True)
    assert cnt_g == cnt_o
    assert cnt_g["paths"] == w * h * spp
    assert img_g.tobytes() == img_o.tobytes(), "max abs diff %g" % np.abs(img_g - img_o).max()
    assert np.array_equal(img8_g, img8_o)


def test_forward_vertex_normals_and_bvh(ipt, oracle):
    """sphere.obj carries per-vertex normals (Triangle::getNormal path) and pushes the scene into BVH mode"""
    objects = scene_objects("cornell_sphere")
    g, o = _pair(ipt, oracle, objects)
    assert g.uses_bvh
    pg = ipt.default_params(width=48, height=48, spp=8, max_depth=6, seed=9)
    po = oracle.make_params(48, 48, 8, max_depth=6, seed=9)
    img_g, cnt_g = g.render(pg)
    img_o, cnt_o = o.render(po)
    assert cnt_g == cnt_o
    assert img_g.tobytes() == img_o.tobytes()


@pytest.mark.parametrize("scene,w,h,spp,depth,gs,slots", [("cornell_sphere", 48, 40, 70, 6, 32, 0), ("cornell_sphere", 48, 40, 70, 0, 32, 5000),
                                                          ("cornell_sphere", 37, 29, 9, 5, 4, 700), ("scene0", 64, 48, 16, 0, 64, 1000),
                                                          ("rotated_cubes", 48, 48, 8, 8, 3, 100000)])
def test_wavefront_integrator_bit_exact(ipt, oracle, scene, w, h, spp, depth, gs, slots):
    """BVH scenes render through the wavefront integrator (state in HBM, one closest-hit kernel per bounce);
    it must reproduce the oracle -- and therefore the megakernel -- bit for bit, whatever the batch size
    (slots < pixels forces pixel tiles, slots < pixels * group forces sample chunks) and with interleaved shards."""
    objects = _scene0(oracle) if scene == "scene0" else scene_objects(scene)
    g = ipt.Scene(objects, asset_root=ASSETS, traversal_mode=2)
    o = oracle.OracleScene(objects, ASSETS)
    assert g.uses_bvh
    pg = ipt.default_params(width=w, height=h, spp=spp, max_depth=depth, seed=77, group_samples=gs)
    po = oracle.make_params(w, h, spp, max_depth=depth, seed=77, group_samples=gs)
    img_o, cnt_o = o.render(po)
    g.set_integrator(2, slots)
    img_w, cnt_w = g.render(pg)
    shard = ipt.default_params(width=w, height=h, spp=spp, max_depth=depth, seed=77, group_samples=gs, row_begin=1, row_end=h, row_step=3)
    img_s, _ = g.render(shard)
    g.set_integrator(1, 0)
    img_m, cnt_m = g.render(pg)
    assert cnt_w == cnt_o and cnt_m == cnt_o
    assert img_w.tobytes() == img_o.tobytes(), "wavefront vs oracle: max abs diff %g" % np.abs(img_w - img_o).max()
    assert img_m.tobytes() == img_o.tobytes()
    assert img_s.tobytes() == img_o[1::3].tobytes()


def test_wavefront_on_bruteforce_scene_equals_oracle(ipt, oracle):
    """the wavefront integrator can also drive a brute-force (shared-memory) scene -- Scene.set_integrator(2), used for
    the wavefront-vs-megakernel A/B on the Cornell box (profiles/r02_cornell_wavefront_ab.txt); same image, bit for bit"""
    objects = _scene0(oracle)
    g = ipt.Scene(objects, asset_root=ASSETS)
    o = oracle.OracleScene(objects, ASSETS)
    assert not g.uses_bvh
    img_o, cnt_o = o.render(oracle.make_params(48, 40, 8, max_depth=6, seed=9))
    for integ in (2, 0):
        g.set_integrator(integ, 0)
        img_g, cnt_g = g.render(ipt.default_params(width=48, height=40, spp=8, max_depth=6, seed=9))
        assert cnt_g == cnt_o and img_g.tobytes() == img_o.tobytes(), integ


def test_large_triangle_set_does_not_change_results(ipt, oracle):
    """BVH scenes keep their large triangles (here the 18 Cornell walls around the 1280-triangle sphere) out of the
    tree and test them brute force first; hits and images must equal the oracle with the set on, capped and off"""
    objects = scene_objects("cornell_sphere")
    o = oracle.OracleScene(objects, ASSETS)
    rays = np.concatenate([random_rays(20000, 5), adversarial_rays(o)])
    tri_o, t_o = o.trace(rays)
    po = oracle.make_params(40, 32, 6, max_depth=5, seed=4)
    img_o, cnt_o = o.render(po)
    sizes = []
    try:
        for limit in (128, 7, 0):
            ipt.set_bvh_large_triangle_limit(limit)
            g = ipt.Scene(objects, asset_root=ASSETS)
            assert g.uses_bvh
            sizes.append(g.n_large)
            tri_g, t_g = g.trace(rays)
            assert np.array_equal(tri_g, tri_o) and t_g.tobytes() == t_o.tobytes(), "limit %d" % limit
            for integ in (1, 2):
                g.set_integrator(integ, 0)
                img_g, cnt_g = g.render(ipt.default_params(width=40, height=32, spp=6, max_depth=5, seed=4))
                assert cnt_g == cnt_o and img_g.tobytes() == img_o.tobytes(), "limit %d integrator %d" % (limit, integ)
            g.close()
    finally:
        ipt.set_bvh_large_triangle_limit(128)
    assert sizes == [18, 7, 0]


def test_emissive_mesh_many_lights(ipt, oracle, tmp_path):
    """an emissive mesh: 1280 + 2 lights.  The light pick becomes a bisection over the cdf (same index as the reference's
    linear scan) and the light table no longer fits the megakernel's shared-memory budget (global-memory path)."""
    src = open(os.path.join(ASSETS, "shapes", "sphere.obj")).read()
    obj, mtl = tmp_path / "glow.obj", tmp_path / "glow.mtl"
    obj.write_text("mtllib glow.mtl\nusemtl glow\n" + src)
    mtl.write_text("newmtl glow\nKd 0.6 0.5 0.4\nKe 1.5 1.2 0.9\n")
    objects = scene_objects("cornell") + [{"pos": [0.3, -1.0, 4.0], "ori": [0, 0.3, 0], "scl": [1.2, 1.2, 1.2], "obj": str(obj), "mtl": str(mtl)}]
    g, o = _pair(ipt, oracle, objects)
    assert g.uses_bvh and g.nE == o.nE == 1282
    po = oracle.make_params(40, 32, 6, max_depth=5, seed=21)
    img_o, cnt_o = o.render(po)
    for integ in (1, 2):
        g.set_integrator(integ, 0)
        img_g, cnt_g = g.render(ipt.default_params(width=40, height=32, spp=6, max_depth=5, seed=21))
        assert cnt_g == cnt_o, integ
        assert img_g.tobytes() == img_o.tobytes(), "integrator %d: max abs diff %g" % (integ, np.abs(img_g - img_o).max())
    # the adjoint (megakernel, lights in global memory) against the oracle's
    target = np.random.RandomState(2).randint(0, 256, (32, 40, 3)).astype(np.uint8)
    loss_g, grad_g, _ = g.render_grad(ipt.default_params(width=40, height=32, spp=6, max_depth=5, seed=21), target)
    loss_o, grad_o = o.render_grad(po, target)[:2]
    assert abs(loss_g - loss_o) <= 1e-6 * abs(loss_o)
    scale = np.abs(grad_o).max()
    assert np.abs(grad_g - grad_o).max() <= 1e-3 * scale


def test_emissive_mesh_32_bit_light_pick(ipt, oracle, tmp_path):
    """more than 4096 lights (a 5120-triangle emissive sphere): the light pick switches from the 16-bit field of the bounce
    word to a full 32-bit uniform of its own Philox call (IPT_STREAM_PICK), in both integrators and in the oracle"""
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("stress_config5", os.path.join(root, "tools", "stress_config5.py"))
    stress = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(stress)
    v, f = stress.read_obj(os.path.join(ASSETS, "shapes", "sphere.obj"))
    v, f = stress.subdivide(v, f, 0.5)
    obj, mtl = tmp_path / "glow.obj", tmp_path / "glow.mtl"
    stress.write_obj(str(obj) + ".body", v, f, normals=False)
    obj.write_text("mtllib glow.mtl\nusemtl glow\n" + open(str(obj) + ".body").read())
    mtl.write_text("newmtl glow\nKd 0.6 0.5 0.4\nKe 1.5 1.2 0.9\n")
    objects = scene_objects("cornell") + [{"pos": [0.3, -1.0, 4.0], "ori": [0, 0.3, 0], "scl": [1.2, 1.2, 1.2], "obj": str(obj), "mtl": str(mtl)}]
    g, o = _pair(ipt, oracle, objects)
    assert g.uses_bvh and g.nE == o.nE == 5122
    img_o, cnt_o = o.render(oracle.make_params(24, 16, 4, max_depth=4, seed=2), threads=oracle.hardware_threads())
    for integ in (1, 2):
        g.set_integrator(integ, 0)
        img_g, cnt_g = g.render(ipt.default_params(width=24, height=16, spp=4, max_depth=4, seed=2))
        assert cnt_g == cnt_o and img_g.tobytes() == img_o.tobytes(), integ


def test_row_sharding_and_determinism(ipt, oracle):
    g = ipt.Scene(_scene0(oracle), asset_root=ASSETS)
    p = ipt.default_params(width=96, height=80, spp=32, seed=3)
    full, _ = g.render(p)
    again, _ = g.render(p)
    assert full.tobytes() == again.tobytes()
    parts = []
    for rb, re in ((0, 11), (11, 40), (40, 80)):
        q = ipt.default_params(width=96, height=80, spp=32, seed=3, row_begin=rb, row_end=re)
        parts.append(g.render(q)[0])
    assert np.concatenate(parts, 0).tobytes() == full.tobytes()
    # interleaved shards (rows r, r+3, ...): what bench.py uses at N > 1 for load balance
    inter = np.empty_like(full)
    for r0 in range(3):
        q = ipt.default_params(width=96, height=80, spp=32, seed=3, row_begin=r0, row_end=80, row_step=3)
        inter[r0::3] = g.render(q)[0]
    assert inter.tobytes() == full.tobytes()


def test_golden_image_statistics(ipt, oracle):
    """the reference's only known-answer output: preds/0_true.png (500x500, 100 spp)"""
    G = json.load(open(os.path.join(GOLDEN, "golden_0_true.json")))
    from PIL import Image
    gold = np.asarray(Image.open(os.path.join(GOLDEN, "golden_0_true.png")).convert("RGB")).astype(np.float32)
    g = ipt.Scene(_scene0(oracle), asset_root=ASSETS)
    p = ipt.default_params(seed=2024)
    _, img8, cnt = g.render(p, want_u8=True)
    for k, (r0, r1, c0, c1) in G["regions"].items():
        m = img8[r0:r1, c0:c1].reshape(-1, 3).mean(0)
        assert np.abs(m - np.array(G["region_means_u8"][k])).max() < 1.5, (k, m, G["region_means_u8"][k])
    rmse = np.sqrt((((img8.astype(np.float32) - gold) / 255.0) ** 2).mean())
    assert rmse < 0.03, rmse   # two independent reference renders differ by 0.019


def test_graph_parity(ipt, oracle):
    objects = _scene0(oracle)
    g, o = _pair(ipt, oracle, objects)
    w, h, spp = 64, 64, 24
    rng = np.random.RandomState(0)
    target = rng.randint(0, 256, (h, w, 3)).astype(np.uint8)
    pg = ipt.default_params(width=w, height=h, spp=spp, seed=77)
    po = oracle.make_params(w, h, spp, seed=77)
    raw_g, data_g, cnt_g = g.create_graph(pg, target)
    raw_o, data_o, cnt_o = o.create_graph(po, target)
    assert cnt_g == cnt_o
    assert np.array_equal(raw_g[:, 0], raw_o[:, 0])                       # update counts are exact
    assert raw_g[-g.nT:, 0].sum() == cnt_g["paths"]                        # eye row: one update per primary hit (closed box)
    np.testing.assert_allclose(raw_g, raw_o, rtol=2e-4, atol=1e-6)
    np.testing.assert_allclose(data_g, data_o, rtol=1e-3, atol=1e-6)
    size = (g.nT + 1) * g.nT
    rows = data_g[:size].reshape(g.nT + 1, g.nT).sum(1)
    assert np.all((np.abs(rows - 1) < 1e-4) | (rows == 0))
    assert not np.isnan(data_g).any()


def test_grad_parity_and_finite_differences(ipt, oracle):
    objects = _scene0(oracle)
    g, o = _pair(ipt, oracle, objects)
    w, h, spp = 32, 32, 16
    rng = np.random.RandomState(1)
    target = rng.randint(0, 256, (h, w, 3)).astype(np.uint8)
    kd = rng.uniform(0.2, 0.8, (g.nT, 3)).astype(np.float32)
    g.set_materials(kd)
    o.set_materials(kd)
    pg = ipt.default_params(width=w, height=h, spp=spp, seed=5, max_depth=64)
    po = oracle.make_params(w, h, spp, seed=5, max_depth=64)
    loss_g, grad_g, img_g = g.render_grad(pg, target)
    loss_o, grad_o, img_o = o.render_grad(po, target)
    assert img_g.tobytes() == img_o.tobytes()
    # interleaved row shards add up to the whole-image loss and gradient (the all-reduce of the inverse pass)
    ls, gs = 0.0, np.zeros_like(grad_g)
    for r0 in range(2):
        l_, g_, _ = g.render_grad(ipt.default_params(width=w, height=h, spp=spp, seed=5, max_depth=64, row_begin=r0, row_end=h, row_step=2), target)
        ls += l_
        gs += g_
    assert abs(ls - loss_g) <= 1e-9 * abs(loss_g) and np.abs(gs - grad_g).max() <= 1e-5 * np.abs(grad_g).max()
    assert abs(loss_g - loss_o) <= 1e-9 * abs(loss_o)
    scale = np.abs(grad_o).max()
    assert np.abs(grad_g - grad_o).max() <= 1e-3 * scale
    # central finite differences of the GPU loss itself on a few albedos
    def loss_at(m):
        g.set_materials(m)
        return g.render_grad(pg, target)[0]
    order = np.argsort(-np.abs(grad_o).ravel())[:4]
    for flat in order:
        t, c = divmod(int(flat), 3)
        e = 1e-2
        mp, mm = kd.copy(), kd.copy()
        mp[t, c] += e
        mm[t, c] -= e
        fd = (loss_at(mp) - loss_at(mm)) / (2 * e)
        assert abs(fd - grad_g[t, c]) <= 2e-3 * abs(fd) + 1e-9, (t, c, fd, grad_g[t, c])


def test_legacy_abi_roundtrip(ipt, oracle, tmp_path, monkeypatch):
    """the six reference entry points through the host-side mirror of ipt_cuda.py"""
    from inverse_path_tracer_b200 import ipt_cuda
    from PIL import Image
    monkeypatch.setenv("IPT_SEED", "42")
    monkeypatch.setattr(ipt_cuda, "ASSET_ROOT", ASSETS)
    scenefile = os.path.join(ASSETS, "scenes", "0.txt")
    n_t, ptr = ipt_cuda.load_scene(scenefile)
    assert n_t == 30
    L = ipt.lib()
    png = str(tmp_path / "img.png")
    L.createImage(ptr, png.encode())
    im = np.asarray(Image.open(png))
    assert im.shape == (500, 500, 3) and im.dtype == np.uint8
    o = oracle.OracleScene.from_file(scenefile, ASSETS)
    _, ref8, _ = o.render(oracle.make_params(500, 500, 100, seed=42, row_end=40), want_u8=True)
    assert np.array_equal(im[:40], ref8)
    L.freeScene(ptr)
    w, pixel, light, labels = ipt_cuda.generate_data(scenefile, png)
    assert w.shape == (31, 30) and pixel.shape == (31, 30, 3) and light.shape == (31, 30, 3) and labels.shape == (30, 3)
    np.testing.assert_allclose(labels, o.get_materials(), rtol=0, atol=0)
    out = str(tmp_path / "pred.png")
    ipt_cuda.render_with_materials(scenefile, out, np.full((30, 3), 0.5, np.float32))
    assert np.asarray(Image.open(out)).shape == (500, 500, 3)


@pytest.mark.parametrize("n,seed", [(1, 0), (2, 1), (5, 2), (9, 3), (64, 4), (1000, 5), (4097, 6), (50000, 7), (1000000, 8)])
def test_gpu_bvh_builder_layout_bit_exact(ipt, oracle, n, seed):
    """the CUDA builder (level-synchronous, closed-form Lomuto permutation) against the oracle's sequential bvh.h restatement"""
    from inverse_path_tracer_b200 import native
    rng = np.random.RandomState(seed)
    pos = rng.uniform(-3, 3, (n, 3)).astype(np.float32)
    ext = rng.uniform(0, 0.3, (n, 3)).astype(np.float32)
    boxes = np.concatenate([pos - ext, pos + ext], 1).astype(np.float32)
    nb_g, ni_g, perm_g = native.build_bvh(boxes, pos, builder=2)
    nb_o, ni_o, perm_o = oracle.bvh_build(boxes, pos)
    assert np.array_equal(perm_g, perm_o) and np.array_equal(ni_g, ni_o) and nb_g.tobytes() == nb_o.tobytes()


def test_gpu_bvh_builder_degenerate_and_scene_inputs(ipt, oracle):
    from inverse_path_tracer_b200 import native
    rng = np.random.RandomState(9)
    grid = np.stack(np.meshgrid(np.arange(16), np.arange(15), np.arange(14), indexing="ij"), -1).reshape(-1, 3).astype(np.float32)
    dup = rng.uniform(-1, 1, (700, 3)).astype(np.float32)
    line = np.zeros((5000, 3), np.float32)
    line[:, 1] = np.repeat(np.arange(100), 50)
    zeros = np.array([[0.0, -0.0, 0.0], [-0.0, 0.0, 1.0], [0.0, 0.0, -0.0], [-0.0, -0.0, 2.0], [1.0, 0.0, 0.0], [-0.0, 1.0, -0.0]], np.float32)
    cases = [np.tile(np.array([[1.0, 2.0, 3.0]], np.float32), (3700, 1)), grid, rng.permutation(grid), line, np.concatenate([dup, dup, dup[:70]]), zeros]
    o = oracle.OracleScene(scene_objects("cornell_sphere"), ASSETS)
    t = o.triangle_table()
    for pos in cases + [t["center"]]:
        boxes = np.concatenate([pos - 0.01, pos + 0.02], 1).astype(np.float32) if pos is not t["center"] else np.concatenate([t["verts"].min(1), t["verts"].max(1)], 1)
        a, b = native.build_bvh(boxes, pos, builder=2), oracle.bvh_build(boxes, pos)
        assert np.array_equal(a[2], b[2]) and np.array_equal(a[1], b[1]) and a[0].tobytes() == b[0].tobytes()


def test_grad_parity_with_phong_lobe(ipt, oracle):
    """the adjoint on a scene with Ks != 0 (Cornell box + loader-fuzz mesh with specular materials and vertex normals):
    loss equal, gradient within 1e-3 of the oracle's double-precision adjoint, which tests/test_oracle.py holds to
    central finite differences on the same kind of scene; both integrator templates (brute force and BVH)"""
    objs = [{"pos": [0, 0, 5], "ori": [0.1, 0.2, 0.3], "scl": [1, 1, 1], "obj": "./fuzz/case_18.obj", "mtl": "./fuzz/case_18.mtl"}] + scene_objects("cornell")
    o = oracle.OracleScene(objs, ASSETS)
    assert np.abs(o.triangle_table()["ks"]).max() > 0.5
    rng = np.random.RandomState(6)
    kd = rng.uniform(0.2, 0.8, (o.nT, 3)).astype(np.float32)
    o.set_materials(kd)
    w, h, spp = 24, 24, 16
    target = rng.randint(0, 256, (h, w, 3)).astype(np.uint8)
    loss_o, grad_o, img_o = o.render_grad(oracle.make_params(w, h, spp, seed=3, max_depth=64), target, threads=oracle.hardware_threads())
    for mode, integ in ((0, 0), (2, 1), (2, 2)):   # brute force; BVH through the megakernel; BVH through the wavefront (recorded vertices + k_wf_sweep)
        g = ipt.Scene(objs, asset_root=ASSETS, traversal_mode=mode)
        g.set_integrator(integ, 0)
        g.set_materials(kd)
        loss_g, grad_g, img_g = g.render_grad(ipt.default_params(width=w, height=h, spp=spp, seed=3, max_depth=64), target)
        assert img_g.tobytes() == img_o.tobytes(), (mode, integ)
        assert abs(loss_g - loss_o) <= 1e-9 * abs(loss_o)
        assert np.abs(grad_g - grad_o).max() <= 1e-3 * np.abs(grad_o).max(), (mode, integ, float(np.abs(grad_g - grad_o).max()), float(np.abs(grad_o).max()))
        g.close()


def test_wavefront_adjoint_on_bvh_scene(ipt, oracle):
    """BVH scenes replay their paths for the adjoint through the wavefront integrator (vertex record in HBM, one backward
    sweep per batch) instead of the persistent megakernel; both must give the oracle's loss and gradient, with the record
    forced into several batches as well (plain scene: 32-byte vertices; vertex normals: 64-byte vertices)"""
    for name in ("rotated_cubes", "cornell_sphere"):
        objects = scene_objects(name)
        o = oracle.OracleScene(objects, ASSETS)
        rng = np.random.RandomState(3)
        w, h, spp = 24, 20, 8
        target = rng.randint(0, 256, (h, w, 3)).astype(np.uint8)
        loss_o, grad_o, _ = o.render_grad(oracle.make_params(w, h, spp, seed=4, max_depth=64), target, threads=oracle.hardware_threads())
        g = ipt.Scene(objects, asset_root=ASSETS, traversal_mode=2)
        for integ, slots in ((1, 0), (2, 0), (2, 1000)):
            g.set_integrator(integ, slots)
            loss_g, grad_g, _ = g.render_grad(ipt.default_params(width=w, height=h, spp=spp, seed=4, max_depth=64), target)
            assert abs(loss_g - loss_o) <= 1e-9 * abs(loss_o), (name, integ, slots)
            assert np.abs(grad_g - grad_o).max() <= 1e-3 * np.abs(grad_o).max(), (name, integ, slots, float(np.abs(grad_g - grad_o).max()), float(np.abs(grad_o).max()))
        g.close()


def test_torch_front_end_and_autograd(ipt, oracle):
    """device-pointer entry points through torch tensors on torch's current stream; RenderLoss backward = path adjoint"""
    import torch
    from inverse_path_tracer_b200.torch_api import TorchScene, render_loss
    objects = _scene0(oracle)
    ts = TorchScene(objects, asset_root=ASSETS)
    o = oracle.OracleScene(objects, ASSETS)
    w, h, spp = 48, 40, 8
    p = ipt.default_params(width=w, height=h, spp=spp, seed=21, max_depth=64)
    img = ts.render(p)
    ref, _ = o.render(oracle.make_params(w, h, spp, seed=21, max_depth=64))
    assert img.cpu().numpy().tobytes() == ref.tobytes()
    rng = np.random.RandomState(5)
    target = rng.randint(0, 256, (h, w, 3)).astype(np.uint8)
    kd0 = rng.uniform(0.2, 0.8, (ts.nT, 3)).astype(np.float32)
    kd = torch.tensor(kd0, device="cuda", requires_grad=True)
    loss = render_loss(kd, ts, p, torch.from_numpy(target).cuda())
    loss.backward()
    o.set_materials(kd0)
    loss_o, grad_o, _ = o.render_grad(oracle.make_params(w, h, spp, seed=21, max_depth=64), target)
    assert abs(float(loss.detach()) - loss_o) <= 1e-6 * abs(loss_o)
    assert np.abs(kd.grad.cpu().numpy() - grad_o).max() <= 1e-3 * np.abs(grad_o).max()
    # ray batches on the device + raw edge sums on the device
    rays = torch.from_numpy(random_rays(100000, 3)).cuda()
    tri, t = ts.trace(rays)
    tri_o, t_o = o.trace(rays.cpu().numpy())
    assert np.array_equal(tri.cpu().numpy(), tri_o) and t.cpu().numpy().tobytes() == t_o.tobytes()
    raw = ts.graph_raw(p, torch.from_numpy(target).cuda()).cpu().numpy()
    raw_o, _, _ = o.create_graph(oracle.make_params(w, h, spp, seed=21, max_depth=64), target)
    np.testing.assert_allclose(raw, raw_o, rtol=2e-4, atol=1e-6)


def test_full_size_properties(ipt, oracle):
    """BASELINE config[1] resolution (1024 x 1024, depth 8) at reduced spp: properties that do not need the oracle --
    one path per pixel-sample, run-to-run determinism, row shards reassemble the image bit for bit, different
    sample grouping only reorders a float sum, and the image is statistically the golden Cornell box"""
    g = ipt.Scene(scene_objects("cornell"), asset_root=ASSETS)
    W = H = 1024
    spp = 32
    p = ipt.default_params(width=W, height=H, spp=spp, max_depth=8, seed=77)
    a, ca = g.render(p)
    b, cb = g.render(p)
    assert ca == cb and ca["paths"] == W * H * spp
    assert ca["ext_rays"] <= 8 * ca["paths"] and ca["shadow_rays"] <= ca["ext_rays"]
    assert a.tobytes() == b.tobytes()
    parts = [g.render(ipt.default_params(width=W, height=H, spp=spp, max_depth=8, seed=77, row_begin=rb, row_end=re))[0]
             for rb, re in ((0, 128), (128, 129), (129, 700), (700, 1024))]
    assert np.concatenate(parts, 0).tobytes() == a.tobytes()
    c, _ = g.render(ipt.default_params(width=W, height=H, spp=spp, max_depth=8, seed=77, group_samples=5))
    assert np.allclose(a, c, rtol=1e-5, atol=1e-6) and not np.array_equal(a, c)
    assert np.isfinite(a).all() and (a >= 0).all()
    light = a[215:270, 460:560].mean()       # the light quad seen from the eye: L_e = 10 re-added per bounce (quirk F5)
    assert 14.0 < light < 19.0, light


def test_scene_sharded_batch_smoke(ipt, oracle):
    """BASELINE config 4 in miniature: independent scenes with their own handles render independently"""
    scenes = [ipt.Scene(oracle.parse_scene_file(os.path.join(ASSETS, "scenes", "%d.txt" % i)), asset_root=ASSETS) for i in (3, 4, 5)]
    p = ipt.default_params(width=64, height=64, spp=8, seed=1)
    imgs = [s.render(p)[0] for s in scenes]
    for i, s in zip((3, 4, 5), scenes):
        o = oracle.OracleScene.from_file(os.path.join(ASSETS, "scenes", "%d.txt" % i), ASSETS)
        assert imgs[(3, 4, 5).index(i)].tobytes() == o.render(oracle.make_params(64, 64, 8, seed=1))[0].tobytes()
    assert not np.array_equal(imgs[0], imgs[1])


def test_edge_case_scenes(ipt, oracle, tmp_path):
    """empty scene, a scene without lights, a single triangle, and a one-pixel / one-sample image"""
    p = ipt.default_params(width=16, height=12, spp=4, seed=3)
    po = oracle.make_params(16, 12, 4, seed=3)
    # no objects at all: every ray misses, the image is black, one extension ray per path
    g = ipt.Scene([], asset_root=ASSETS)
    assert g.nT == 0 and g.nE == 0
    img, cnt = g.render(p)
    assert not img.any() and cnt == {"paths": 16 * 12 * 4, "ext_rays": 16 * 12 * 4, "shadow_rays": 0}
    tri, t = g.trace(random_rays(1000, 1))
    assert (tri == -1).all() and np.isinf(t).all()
    # geometry but no emitter (nE == 0): directLighting returns before drawing (path_trace.cu:33-34)
    cube = [{"pos": [0, 0, 3], "ori": [0.3, 0.2, 0.1], "scl": [1, 1, 1], "obj": "./shapes/cube.obj", "mtl": "*Kd 0.5 0.6 0.7*"}]
    g, o = _pair(ipt, oracle, cube)
    assert g.nE == 0
    ig, cg = g.render(p)
    io, co = o.render(po)
    assert cg == co and ig.tobytes() == io.tobytes() and not ig.any() and cg["shadow_rays"] == 0
    raw_g, data_g, _ = g.create_graph(p, np.zeros((12, 16, 3), np.uint8))
    raw_o, data_o, _ = o.create_graph(po, np.zeros((12, 16, 3), np.uint8))
    assert np.array_equal(raw_g[:, 0], raw_o[:, 0]) and not np.isnan(data_g).any()
    # one emissive triangle seen from the eye (an unpaired quad slot, a one-slot group)
    obj = tmp_path / "tri.obj"
    obj.write_text("mtllib t.mtl\nusemtl lamp\nv -1 -1 3\nv 1 -1 3\nv 0 1 3\nf 1 2 3\n")
    (tmp_path / "t.mtl").write_text("newmtl lamp\nKd 0.5 0.5 0.5\nKe 3 2 1\n")
    one = [{"pos": [0, 0, 0], "ori": [0, 0, 0], "scl": [1, 1, 1], "obj": str(obj), "mtl": str(tmp_path / "t.mtl")}]
    g, o = _pair(ipt, oracle, one)
    assert (g.nT, g.nE) == (1, 1)
    ig, cg = g.render(p)
    io, co = o.render(po)
    assert cg == co and ig.tobytes() == io.tobytes() and ig.max() > 0
    # 1 x 1 pixel, 1 sample
    q, qo = ipt.default_params(width=1, height=1, spp=1, seed=9), oracle.make_params(1, 1, 1, seed=9)
    g, o = _pair(ipt, oracle, scene_objects("cornell"))
    assert g.render(q)[0].tobytes() == o.render(qo)[0].tobytes()
    # bad parameters are reported, not executed
    with pytest.raises(RuntimeError, match="positive"):
        g.render(ipt.default_params(width=0, height=4, spp=1))
    with pytest.raises(RuntimeError, match="empty row shard"):
        g.render(ipt.default_params(width=4, height=4, spp=1, row_begin=4, row_end=4))
    with pytest.raises(RuntimeError, match="never terminate"):
        g.render(ipt.default_params(width=4, height=4, spp=1, p_rr=1.0, max_depth=0))
    # p_rr = 1 is fine with a depth limit: every path runs to max_depth or leaves the scene
    a_, ca_ = g.render(ipt.default_params(width=8, height=8, spp=2, p_rr=1.0, max_depth=4, seed=2))
    b_, cb_ = o.render(oracle.make_params(8, 8, 2, p_rr=1.0, max_depth=4, seed=2))
    assert ca_ == cb_ and a_.tobytes() == b_.tobytes() and 3 * ca_["paths"] < ca_["ext_rays"] <= 4 * ca_["paths"]   # a few paths leave through the open front


def test_default_group_rule_matches_oracle(ipt, oracle):
    """both sides resolve group_samples = -1 with the same size rule, and 0 means one running sum per pixel"""
    g, o = _pair(ipt, oracle, scene_objects("cornell"))
    for gs in (-1, 0, 5):
        pg = ipt.default_params(width=40, height=24, spp=12, max_depth=4, seed=8, group_samples=gs)
        po = oracle.make_params(40, 24, 12, max_depth=4, seed=8, group_samples=gs)
        assert g.render(pg)[0].tobytes() == o.render(po)[0].tobytes(), gs
    assert ipt.default_params().group_samples == -1


def _fuzz_scene_cases():
    path = os.path.join(GOLDEN, "loader_fuzz.npz")
    return [str(c) for c in np.load(path)["cases"]] if os.path.exists(path) else []


@pytest.mark.parametrize("chunk", range(4))
def test_fuzz_scenes_bit_exact_all_integrators(ipt, oracle, chunk):
    """The 48 loader-fuzz meshes (random triangle soups: slivers, crossing faces, random Kd/Ks/Ke/Ns materials, vertex
    normals, concave polygons) inside the Cornell box: hits, ray counts and images of the brute-force kernel, the BVH
    megakernel and the wavefront integrator must equal the oracle bit for bit -- including the Phong lobe and specular
    sampling, whose pow() is part of the canon (canon v4)."""
    cases = _fuzz_scene_cases()[chunk::4]
    assert cases
    cornell = scene_objects("cornell")
    rng = np.random.RandomState(7 + chunk)
    rays = np.concatenate([rng.uniform(-2, 2, (3000, 3)) + [0, 0, 4], rng.normal(size=(3000, 3))], 1).astype(np.float32)
    rays[:, 3:] /= np.linalg.norm(rays[:, 3:], axis=1, keepdims=True)
    for case in cases:
        objs = [{"pos": [0, 0, 5], "ori": [0.1, 0.2, 0.3], "scl": [1, 1, 1], "obj": "./fuzz/%s.obj" % case, "mtl": "./fuzz/%s.mtl" % case}] + cornell
        o = oracle.OracleScene(objs, ASSETS)
        img_o, cnt_o = o.render(oracle.make_params(32, 24, 4, max_depth=5, seed=3))
        tri_o, t_o = o.trace(rays)
        for mode in (0, 2):
            g = ipt.Scene(objs, asset_root=ASSETS, traversal_mode=mode)
            tri_g, t_g = g.trace(rays)
            assert np.array_equal(tri_g, tri_o) and t_g.tobytes() == t_o.tobytes(), (case, mode)
            for integ in ((1,) if mode == 0 else (1, 2)):
                g.set_integrator(integ, 0)
                img_g, cnt_g = g.render(ipt.default_params(width=32, height=24, spp=4, max_depth=5, seed=3))
                assert cnt_g == cnt_o, (case, mode, integ)
                assert img_g.tobytes() == img_o.tobytes(), (case, mode, integ, float(np.nanmax(np.abs(img_g - img_o))))
            g.close()


def test_graph_data_device_compress_equals_host(ipt, oracle):
    """DataWrapper::compress on the device (torch front end: graph_data) against the host version on the same raw edge
    sums: identical bits (the log of the edge weights is the canon's, not libm's), reference layout, rows sum to 1"""
    import torch
    from inverse_path_tracer_b200 import native
    from inverse_path_tracer_b200.torch_api import TorchScene
    ts = TorchScene(_scene0(oracle), asset_root=ASSETS, device=0)
    target = torch.from_numpy(np.random.RandomState(4).randint(0, 256, (48, 64, 3)).astype(np.uint8)).cuda()
    p = ipt.default_params(width=64, height=48, spp=8, seed=6)
    raw = ts.graph_raw(p, target)
    data = torch.empty(7 * (ts.nT + 1) * ts.nT, dtype=torch.float32, device="cuda")
    import ctypes as C
    native._check(native.lib().iptGraphCompressDevice(C.c_void_p(raw.data_ptr()), ts.nT, C.c_void_p(data.data_ptr()),
                                                      C.c_void_p(torch.cuda.current_stream().cuda_stream)), "iptGraphCompressDevice")
    host = native.graph_compress(raw.cpu().numpy(), ts.nT)
    assert data.cpu().numpy().tobytes() == np.ascontiguousarray(host, np.float32).tobytes()
    w, pixel, light = ts.graph_data(p, target)
    assert w.shape == (ts.nT + 1, ts.nT) and pixel.shape == (ts.nT + 1, ts.nT, 3) and light.shape == (ts.nT + 1, ts.nT, 3)
    sums = w.sum(-1).cpu().numpy()
    assert all(abs(s - 1.0) < 1e-5 or s == 0.0 for s in sums) and sums[-1] > 0.99   # the eye row always sees something
    ts.close()


@pytest.mark.parametrize("w,h,spp,depth,gs,rb,re", [(1024, 1024, 1024, 8, 64, 510, 514), (512, 512, 64, 0, -1, 252, 260)])
def test_bench_configs_row_band_bit_exact(ipt, oracle, w, h, spp, depth, gs, rb, re):
    """the bench workloads themselves, not scaled-down stand-ins: a band of rows of BASELINE config 2 (Cornell box
    1024 x 1024 x 1024 spp, depth 8, groups of 64 -- bench.py's headline) and of the metric point (512 x 512 x 64 spp,
    unbounded depth, size-rule groups) against the oracle, bit for bit, ray counters included"""
    g, o = _pair(ipt, oracle, scene_objects("cornell"))
    img_g, cnt_g = g.render(ipt.default_params(width=w, height=h, spp=spp, max_depth=depth, seed=3, group_samples=gs, row_begin=rb, row_end=re))
    img_o, cnt_o = o.render(oracle.make_params(w, h, spp, max_depth=depth, seed=3, group_samples=gs, row_begin=rb, row_end=re), threads=oracle.hardware_threads())
    assert cnt_g == cnt_o
    assert img_g.tobytes() == img_o.tobytes(), float(np.abs(img_g - img_o).max())


def test_reference_ipt_cuda_runs_unmodified(ipt, tmp_path):
    """The reference's OWN ipt_cuda.py (ipt_cuda.py:7-8 dlopens ./build/libipt.so and ./build/libpt.so), byte for byte,
    against this library: generate_files(1) (scene file + createImage), generate_data (createGraph + getMaterials) and
    render_with_materials (setMaterials + createImage), ipt_cuda.py:115-183.  The file is taken from the read-only
    reference tree (or from $IPT_REFERENCE_IPT_CUDA) at test time and is never part of this repository; skipped where
    neither exists."""
    import shutil
    import subprocess
    import sys
    src = os.environ.get("IPT_REFERENCE_IPT_CUDA") or "/root/reference/ipt_cuda.py"
    if not os.path.exists(src):
        pytest.skip("reference ipt_cuda.py not available here")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    shutil.copyfile(src, tmp_path / "ipt_cuda.py")
    os.symlink(os.path.join(root, "build"), tmp_path / "build")
    os.symlink(os.path.join(ASSETS, "CornellBox"), tmp_path / "CornellBox")
    os.symlink(os.path.join(ASSETS, "shapes"), tmp_path / "shapes")
    os.mkdir(tmp_path / "scenes")
    os.mkdir(tmp_path / "imgs")
    driver = (
        "import numpy as np, torch\n"
        "np.random.seed(4)\n"
        "import ipt_cuda\n"
        "ipt_cuda.generate_files(1)\n"
        "w, pixel, light, labels = ipt_cuda.generate_data('scenes/0.txt', 'imgs/0.png')\n"
        "ipt_cuda.render_with_materials('scenes/0.txt', 'imgs/0_pred.png', torch.from_numpy(labels))\n"
        "np.savez('out.npz', w=w, pixel=pixel, light=light, labels=labels)\n")
    env = dict(os.environ, IPT_SEED="123")
    r = subprocess.run([sys.executable, "-c", driver], cwd=tmp_path, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    from PIL import Image
    a = np.asarray(Image.open(tmp_path / "imgs" / "0.png").convert("RGB")).astype(np.float32) / 255.0
    b = np.asarray(Image.open(tmp_path / "imgs" / "0_pred.png").convert("RGB")).astype(np.float32) / 255.0
    assert a.shape == (500, 500, 3) and b.shape == (500, 500, 3)
    assert np.sqrt(((a - b) ** 2).mean()) < 1e-6          # same materials, same IPT_SEED: the same image
    z = np.load(tmp_path / "out.npz")
    nT = 30
    assert z["w"].shape == (nT + 1, nT) and z["pixel"].shape == (nT + 1, nT, 3) and z["light"].shape == (nT + 1, nT, 3) and z["labels"].shape == (nT, 3)
    rows = z["w"].sum(1)
    assert np.all((np.abs(rows - 1.0) < 1e-4) | (rows == 0.0)) and rows[nT] > 0.99      # row-normalised transport weights, eye row last
    assert not np.isnan(z["pixel"]).any()
    assert np.allclose(z["labels"][0], [0.725, 0.71, 0.68], atol=1e-6)                  # floor of the Cornell box (CornellBox-Empty-CO.mtl)
    # the golden-image statistics of the reference's own render hold for the image its unmodified front end produced
    G = json.load(open(os.path.join(GOLDEN, "golden_0_true.json")))
    img8 = (a * 255.0 + 0.5).astype(np.uint8)
    # (generate_files gives the cube a random albedo -- np.random.seed above -- so only regions it barely influences)
    for k, tol in (("light", 1.5), ("back_wall", 8.0), ("ceiling", 8.0)):
        r0, r1, c0, c1 = G["regions"][k]
        m = img8[r0:r1, c0:c1].reshape(-1, 3).mean(0)
        assert np.abs(m - np.array(G["region_means_u8"][k])).max() < tol, (k, m.tolist(), G["region_means_u8"][k])

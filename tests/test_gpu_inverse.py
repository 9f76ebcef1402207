"""The in-library inverse loop (include/ipt_b200.h part 3, csrc/inverse.cu) against the Python loop it replaces
(torch_api.render_loss + torch.optim.Adam), and the fused exchange step with two ranks.

The reference pins nothing here (its material recovery is a GCN, ipt.py); the checks are (a) the library step equals
loss_and_grad + torch's Adam on the same seeds, (b) two ranks that each render half of the rows and exchange their
partial sums through the mailbox protocol land on the albedos of one rank rendering everything."""
import os

import numpy as np
import pytest

from conftest import ASSETS, GOLDEN

pytestmark = pytest.mark.gpu


def _scene0_objects():
    from inverse_path_tracer_b200.ipt_cuda import parse_scene_file
    return parse_scene_file(os.path.join(ASSETS, "scenes", "0.txt"))


def _target(torch, w, h, seed=5):
    rng = np.random.RandomState(seed)
    return torch.from_numpy(rng.randint(0, 256, (h, w, 3)).astype(np.uint8)).cuda()


def test_inverse_step_equals_python_adam_loop(ipt):
    import torch
    from inverse_path_tracer_b200.torch_api import TorchScene, inverse_optimizer, render_loss
    W, H, SPP, STEPS = 64, 48, 8, 6
    target = _target(torch, W, H)
    a = TorchScene(_scene0_objects(), asset_root=ASSETS)
    b = TorchScene(_scene0_objects(), asset_root=ASSETS)
    kd0 = np.full((a.nT, 3), 0.5, np.float32)
    # library loop
    p = ipt.default_params(width=W, height=H, spp=SPP, max_depth=0, seed=0)
    opt = inverse_optimizer(a, p, target, lr=0.02)
    opt.set_kd(kd0)
    opt.step(100, STEPS)
    loss_lib, kd_lib, steps, hist = opt.state(history=STEPS)
    assert steps == STEPS
    # Python loop on a second handle of the same scene
    kd = torch.full((b.nT, 3), 0.5, device="cuda", requires_grad=True)
    adam = torch.optim.Adam([kd], lr=0.02)
    losses = []
    for i in range(STEPS):
        adam.zero_grad(set_to_none=True)
        loss = render_loss(kd, b, ipt.default_params(width=W, height=H, spp=SPP, max_depth=0, seed=100 + i), target, reduce=False)
        loss.backward()
        adam.step()
        with torch.no_grad():
            kd.clamp_(1e-3, 1.0)
        losses.append(float(loss.detach()))
    kd_py = kd.detach().cpu().numpy()
    # the first loss is taken at identical albedos: same kernels, same sums (render_loss returns it as float32)
    assert abs(hist[0] - losses[0]) <= 1e-6 * abs(losses[0])
    assert np.abs(kd_lib - kd_py).max() < 2e-5, np.abs(kd_lib - kd_py).max()
    assert abs(loss_lib - losses[-1]) <= 1e-4 * abs(losses[-1])
    # getMaterials reflects the device-side update (the reference's getMaterials always returns the last set)
    assert np.abs(a.scene.get_materials() - kd_lib).max() == 0.0
    opt.close()
    a.close()
    b.close()


def test_inverse_step_on_bvh_scene_equals_python_adam_loop(ipt):
    """the same equality on a scene that runs through the BVH / wavefront path (Cornell box + sphere mesh with vertex normals,
    1298 triangles): iptInverseStep's BVH branch (forward without resolve, k_resolve_loss, wavefront adjoint) against render_loss +
    torch Adam on a second handle"""
    import torch
    from conftest import scene_objects
    from inverse_path_tracer_b200.torch_api import TorchScene, inverse_optimizer, render_loss
    W, H, SPP, STEPS = 48, 40, 8, 4
    target = _target(torch, W, H, seed=2)
    a = TorchScene(scene_objects("cornell_sphere"), asset_root=ASSETS)
    b = TorchScene(scene_objects("cornell_sphere"), asset_root=ASSETS)
    assert a.scene.uses_bvh
    kd0 = np.full((a.nT, 3), 0.6, np.float32)
    opt = inverse_optimizer(a, ipt.default_params(width=W, height=H, spp=SPP, max_depth=6, seed=0), target, lr=0.02)
    opt.set_kd(kd0)
    opt.step(40, STEPS)
    loss_lib, kd_lib, steps, hist = opt.state(history=STEPS)
    assert steps == STEPS
    kd = torch.full((b.nT, 3), 0.6, device="cuda", requires_grad=True)
    adam = torch.optim.Adam([kd], lr=0.02)
    losses = []
    for i in range(STEPS):
        adam.zero_grad(set_to_none=True)
        loss = render_loss(kd, b, ipt.default_params(width=W, height=H, spp=SPP, max_depth=6, seed=40 + i), target, reduce=False)
        loss.backward()
        adam.step()
        with torch.no_grad():
            kd.clamp_(1e-3, 1.0)
        losses.append(float(loss.detach()))
    assert abs(hist[0] - losses[0]) <= 1e-6 * abs(losses[0])
    assert np.abs(kd_lib - kd.detach().cpu().numpy()).max() < 2e-5
    assert abs(loss_lib - losses[-1]) <= 1e-4 * abs(losses[-1])
    opt.close()
    a.close()
    b.close()


def test_inverse_two_ranks_exchange_on_one_gpu(ipt):
    """two optimisers (rank 0 / 1 of world 2, interleaved row shards) in one process on one GPU, mailboxes connected by
    raw pointers: the exchange protocol of k_inverse_update must deliver the same albedos to both and match one rank
    rendering the whole image (sum order differs: 1e-6)"""
    import torch
    from inverse_path_tracer_b200.torch_api import TorchScene
    from inverse_path_tracer_b200 import native
    W, H, SPP, STEPS = 64, 48, 8, 5
    target = _target(torch, W, H, seed=8)
    kd0 = np.full((30, 3), 0.4, np.float32)
    scenes = [TorchScene(_scene0_objects(), asset_root=ASSETS) for _ in range(3)]
    full = native.InverseOptimizer(scenes[2].scene, ipt.default_params(width=W, height=H, spp=SPP, max_depth=0), target.data_ptr())
    full.set_kd(kd0)
    full.step(7, STEPS)
    loss_f, kd_f, _ = full.state()
    ranks = []
    for r in range(2):
        p = ipt.default_params(width=W, height=H, spp=SPP, max_depth=0, row_begin=r, row_end=H, row_step=2)
        o = native.InverseOptimizer(scenes[r].scene, p, target.data_ptr(), rank=r, world=2, timeout_ms=20000)
        o.set_kd(kd0)
        ranks.append(o)
    boxes = [o.mailbox() for o in ranks]
    for o in ranks:
        o.connect_local(boxes)
    for i in range(STEPS):          # every call is asynchronous; rank 0's update kernel waits on the device for rank 1's
        for o in ranks:
            o.step(7 + i, 1)
    (l0, k0, s0), (l1, k1, s1) = ranks[0].state(), ranks[1].state()
    assert s0 == s1 == STEPS
    assert l0 == l1 and k0.tobytes() == k1.tobytes(), "ranks must hold bit-identical sums and albedos"
    assert abs(l0 - loss_f) <= 1e-6 * abs(loss_f)
    assert np.abs(k0 - kd_f).max() < 1e-5
    for o in ranks + [full]:
        o.close()
    for s in scenes:
        s.close()


def test_inverse_recovers_reference_materials(ipt):
    """BASELINE config 3 end to end, promoted from tools/inverse_config3.py: invert the REFERENCE's own render
    preds/0_true.png (cuRAND, unknown seeds) with the in-library loop from albedo 0.5; the walls, floor and ceiling must
    come back to the reference's materials (getMaterials of scenes/0.txt) within 0.01 on average per surface -- the one reference-generated pin
    on the estimator beyond region means: a forward model that missed a quirk of path_trace.cu converges elsewhere."""
    import torch
    from PIL import Image
    from inverse_path_tracer_b200.torch_api import TorchScene, inverse_optimizer
    target = torch.from_numpy(np.asarray(Image.open(os.path.join(GOLDEN, "golden_0_true.png")).convert("RGB")).copy()).cuda()
    sc = TorchScene(_scene0_objects(), asset_root=ASSETS)
    truth = sc.scene.get_materials().copy()
    opt = inverse_optimizer(sc, ipt.default_params(width=500, height=500, spp=64, max_depth=0), target, lr=0.02)
    opt.set_kd(np.full((sc.nT, 3), 0.5, np.float32))
    opt.step(0, 200)
    loss, kd, steps, hist = opt.state(history=200)
    assert steps == 200 and hist[-1] < 0.1 * hist[0]
    # triangles 0-15: floor, ceiling (8), back wall, right wall, left wall; 16-17 is the light
    # (its albedo is barely observable), 18-29 the cube (faces turned away from the camera are not observable)
    surfaces = {"floor": slice(0, 2), "ceiling": slice(2, 10), "back wall": slice(10, 12), "right wall": slice(12, 14), "left wall": slice(14, 16)}
    for name, sl in surfaces.items():
        err = np.abs(kd[sl] - truth[sl])
        assert err.mean() < 0.01 and err.max() < 0.03, (name, float(err.mean()), float(err.max()))
    opt.close()
    sc.close()


def test_gcn_pipeline_end_to_end(ipt, tmp_path):
    """The reference's learning pipeline on this library (ipt.py:69-140): createGraph data product for a few scenes on
    the device (TorchScene.graph_data), DGL-free GCN (gcn.build_graph / gcn.fit), prediction rendered back with
    setMaterials + forward pass (render_with_materials).  The loss must drop and the predicted image must render."""
    import torch
    from inverse_path_tracer_b200 import gcn
    from inverse_path_tracer_b200.ipt_cuda import parse_scene_file
    from inverse_path_tracer_b200.torch_api import TorchScene
    data, scenes = [], []
    for i in range(4):
        sc = TorchScene(parse_scene_file(os.path.join(ASSETS, "scenes", "%d.txt" % i)), asset_root=ASSETS)
        labels = sc.scene.get_materials().copy()
        tgt = torch.empty((100, 100, 3), dtype=torch.uint8, device="cuda")
        sc.render(ipt.default_params(width=100, height=100, spp=64, seed=i), out_u8=tgt)
        w, pixel, light = sc.graph_data(ipt.default_params(width=100, height=100, spp=64, seed=100 + i), tgt)
        assert w.shape == (sc.nT + 1, sc.nT) and not torch.isnan(pixel).any()
        data.append((w, pixel, light, labels))
        scenes.append(sc)
    _, l_short = gcn.fit(data, epochs=1, lr=1e-3, device="cuda")
    model, l_long = gcn.fit(data, epochs=300, lr=1e-3, device="cuda")
    assert l_long < 0.8 * l_short, (l_short, l_long)
    # render the prediction for scene 0 (ipt_cuda.render_with_materials: setMaterials + createImage)
    W, x = gcn.build_graph(*data[0][:3])
    pred = model(W.cuda(), x.cuda()).detach()
    assert pred.shape == (scenes[0].nT, 3) and float(pred.min()) > 0 and float(pred.max()) < 1
    scenes[0].set_materials(pred)
    img = scenes[0].render(ipt.default_params(width=100, height=100, spp=32, seed=9))
    assert torch.isfinite(img).all() and float(img.mean()) > 0.05
    assert np.abs(scenes[0].scene.get_materials() - pred.cpu().numpy()).max() < 1e-7   # getMaterials sees the device-side set
    for sc in scenes:
        sc.close()


def test_calls_on_two_streams_serialise_on_one_scene(ipt):
    """include/ipt_b200.h "Concurrency": a scene handle runs one call at a time.  Two device-pointer renders issued
    back to back on different CUDA streams share the handle's scratch (work counter, partial sums); the library chains
    them through an event, so both images must equal the ones rendered alone."""
    import torch
    from inverse_path_tracer_b200.torch_api import TorchScene
    sc = TorchScene(_scene0_objects(), asset_root=ASSETS)
    pa = ipt.default_params(width=160, height=120, spp=32, seed=1)
    pb = ipt.default_params(width=160, height=120, spp=32, seed=2)
    ref_a, ref_b = sc.render(pa).clone(), sc.render(pb).clone()
    torch.cuda.synchronize()
    sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
    out_a = torch.empty_like(ref_a)
    out_b = torch.empty_like(ref_b)
    for _ in range(3):
        with torch.cuda.stream(sa):
            sc.render(pa, out=out_a)
        with torch.cuda.stream(sb):
            sc.render(pb, out=out_b)
    torch.cuda.synchronize()
    assert torch.equal(out_a, ref_a) and torch.equal(out_b, ref_b)
    sc.close()

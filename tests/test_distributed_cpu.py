"""world_size-2 gloo tests of the multi-GPU plumbing (inverse_path_tracer_b200/parallel.py) on CPU.

The compute backend is injected: here the CPU oracle renders each rank's row shard (tests may call
the oracle), so the test checks exactly what the GPU path relies on -- pixel-row sharding with a
(pixel, sample, bounce)-keyed RNG reproduces the single-process image bit for bit, and the fused
[loss, grad] buffer all-reduces to the whole-image value."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ASSETS, ROOT, scene_objects


def _worker(rank, world, port, tmpdir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from inverse_path_tracer_b200 import parallel
        from oracle import oracle as O
        sc = O.OracleScene(scene_objects("cornell"), ASSETS)
        W, H, SPP = 24, 19, 6          # 19 rows over 2 ranks: unequal shards

        def render_rows(b, e):
            img, _ = sc.render(O.make_params(W, H, SPP, seed=5, max_depth=6, row_begin=b, row_end=e), threads=2)
            return torch.from_numpy(img)

        full = parallel.render_sharded(render_rows, H)
        # interleaved shards (what bench.py uses): H2 rows, H2 % world == 0
        H2 = 18
        b2, e2, s2 = parallel.shard_rows_interleaved(H2, rank, world)
        part, _ = sc.render(O.make_params(W, H2, SPP, seed=5, max_depth=6, row_begin=b2, row_end=e2, row_step=s2), threads=2)
        inter = parallel.gather_rows_interleaved(torch.from_numpy(part), H2)
        np.save(os.path.join(tmpdir, "inter_%d.npy" % rank), inter.numpy())
        target = np.random.RandomState(3).randint(0, 256, (H, W, 3)).astype(np.uint8)
        b, e = parallel.shard_rows(H, rank, world)
        loss, grad, _ = sc.render_grad(O.make_params(W, H, SPP, seed=5, max_depth=6, row_begin=b, row_end=e), target, threads=2)
        lg = torch.from_numpy(np.concatenate([[loss], grad.ravel()]))
        parallel.allreduce_sum(lg)
        np.save(os.path.join(tmpdir, "img_%d.npy" % rank), full.numpy())
        np.save(os.path.join(tmpdir, "lg_%d.npy" % rank), lg.numpy())
    finally:
        dist.destroy_process_group()


def test_shard_rows_partition():
    from inverse_path_tracer_b200 import parallel
    for h in (1, 7, 500, 1024):
        for w in (1, 2, 3, 8):
            parts = [parallel.shard_rows(h, r, w) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == h
            assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
            sizes = [e - b for b, e in parts]
            assert max(sizes) - min(sizes) <= 1
    assert parallel.scenes_for_rank(100, 3, 8) == list(range(3, 100, 8))


def test_two_rank_gloo_sharded_render_and_gradient(tmp_path, oracle):
    world = 2
    port = 29600 + (os.getpid() % 300)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    sc = oracle.OracleScene(scene_objects("cornell"), ASSETS)
    W, H, SPP = 24, 19, 6
    ref, _ = sc.render(oracle.make_params(W, H, SPP, seed=5, max_depth=6), threads=2)
    target = np.random.RandomState(3).randint(0, 256, (H, W, 3)).astype(np.uint8)
    loss, grad, _ = sc.render_grad(oracle.make_params(W, H, SPP, seed=5, max_depth=6), target, threads=2)
    ref2, _ = sc.render(oracle.make_params(W, 18, SPP, seed=5, max_depth=6), threads=2)
    for r in range(world):
        assert np.load(os.path.join(str(tmp_path), "inter_%d.npy" % r)).tobytes() == ref2.tobytes()
        img = np.load(os.path.join(str(tmp_path), "img_%d.npy" % r))
        assert img.shape == (H, W, 3) and img.tobytes() == ref.tobytes()          # bit-identical for any world size
        lg = np.load(os.path.join(str(tmp_path), "lg_%d.npy" % r))
        assert abs(lg[0] - loss) < 1e-12 and np.abs(lg[1:].reshape(-1, 3) - grad).max() < 1e-12

"""not-gpu: the N > 1 host logic on world_size-2 gloo process groups.

* scene inference: every rank handles its row band (+ halo rows) with the geometry the CUDA
  path uses (vitcnn_b200.utils.band_geometry); the disjoint slices add up to the single-process
  map -- no collective on the data path (the all_reduce here only assembles the test result);
* training: per-rank gradients in the flat canonical bucket, all-reduce(sum) / world equals the
  gradient of the mean of the per-rank losses.
The per-window / per-batch compute is the CPU oracle: this file checks plumbing, not kernels."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.functional as F

from oracle import data_ref as R
from oracle.model_ref import ViTCNNRef


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _window_fn(h, l):
    """Deterministic stand-in for the network: any per-window function works for the plumbing."""
    return np.stack([h.sum((1, 2, 3)), l.sum((1, 2, 3)), h[:, 0].max((1, 2)), h[:, :, 0, 0].mean(1)], 1).astype(np.float32)


def _scene_worker(rank, world, port, H, W, C1, C2, P, stride, out_path):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from vitcnn_b200.utils import band_geometry
        img1, img2, _ = R.synthetic_scene(H, W, C1, C2, 5, seed=7)
        geo = band_geometry(H, W, P, stride, rank, world)
        full = np.zeros((H, W, 4), dtype=np.float64)
        if geo["count"]:
            band1, band2 = img1[geo["x0"]:geo["x1"]], img2[geo["x0"]:geo["x1"]]
            probs = R.scene_test(_window_fn, band1, band2, P, 64, 4, stride)
            # band-local window rows are geo["xs"]; the band owns map rows o0:o1
            full[geo["o0"]:geo["o1"]] = probs[P // 2:P // 2 + geo["o1"] - geo["o0"]]
        t = torch.from_numpy(full)
        dist.all_reduce(t)
        if rank == 0:
            np.save(out_path, t.numpy())
    finally:
        dist.destroy_process_group()


def test_row_bands_two_ranks(tmp_path):
    H, W, C1, C2, P = 29, 37, 6, 1, 7
    for stride in (1, 3):
        out = str(tmp_path / f"scene{stride}.npy")
        mp.spawn(_scene_worker, args=(2, _free_port(), H, W, C1, C2, P, stride, out), nprocs=2, join=True)
        img1, img2, _ = R.synthetic_scene(H, W, C1, C2, 5, seed=7)
        want = R.scene_test(_window_fn, img1, img2, P, 64, 4, stride)
        got = np.load(out)
        if stride == 1:
            assert np.array_equal(got, want)
        else:   # clamped last window row may be owned by one band only: maps still agree
            assert np.array_equal(got, want)


def _make_model():
    torch.manual_seed(0)
    return ViTCNNRef(12, 1, patch_size=5, num_classes=4, dropout=0.0).train()


def _batch():
    g = torch.Generator().manual_seed(3)
    return (torch.rand(10, 12, 5, 5, generator=g), torch.rand(10, 1, 5, 5, generator=g),
            torch.randint(1, 4, (10,), generator=g))


def _train_worker(rank, world, port, out_path):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from vitcnn_b200.train import allreduce_mean_, flatten_grads, shard_batch
        model = _make_model()
        hsi, lid, y = _batch()
        sl = shard_batch(len(y), rank, world)
        w = torch.ones(4)
        w[0] = 0
        F.cross_entropy(model(hsi[sl], lid[sl]), y[sl], weight=w).backward()
        flat = allreduce_mean_(flatten_grads(model))
        if rank == 0:
            np.save(out_path, flat.numpy())
    finally:
        dist.destroy_process_group()


def test_flat_gradient_bucket_allreduce_two_ranks(tmp_path):
    from vitcnn_b200.train import flat_offsets, flatten_grads, param_names, shard_batch
    out = str(tmp_path / "grads.npy")
    mp.spawn(_train_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    model = _make_model()
    hsi, lid, y = _batch()
    w = torch.ones(4)
    w[0] = 0
    loss = 0
    for r in range(2):      # mean of the per-rank losses (each rank: its own BatchNorm statistics)
        sl = shard_batch(len(y), r, 2)
        loss = loss + F.cross_entropy(model(hsi[sl], lid[sl]), y[sl], weight=w) / 2
    loss.backward()
    want = flatten_grads(model).numpy()
    got = np.load(out)
    assert got.shape == want.shape
    assert np.abs(got - want).max() <= 1e-6 * max(1.0, np.abs(want).max())
    names = param_names()
    assert len(names) == 58 and len(set(names)) == 58 and set(names) == set(dict(model.named_parameters()))
    offs, total = flat_offsets([p.numel() for p in model.parameters()])
    assert all(o % 4 == 0 for o in offs) and total == got.shape[0]
    assert [shard_batch(10, r, 3) for r in range(3)] == [slice(0, 4), slice(4, 7), slice(7, 10)]

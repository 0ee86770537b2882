"""-m gpu: paths added in round 2 -- streaming host API, tensor-map TMA gather on awkward channel counts, the
two-threads-per-row token kernel, the per-variant fallback of the shared stem (environment switches are read once per
process, so those run in a child interpreter)."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import data_ref as R
from tests.test_gpu_model import DEV, make_pair

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_streaming_host_api_equals_the_synchronous_one():
    """predict_scene_host(sync=False), scene after scene into double-buffered pinned results, must deliver the maps the
    synchronous call delivers (which equal predict_scene's: tests/test_gpu_scene.py)."""
    import vitcnn_b200
    H, W, C1, C2, P, K = 97, 64, 32, 1, 11, 6
    _, net = make_pair(C1, C2, P, K, seed=5)
    scenes = [R.synthetic_scene(H, W, C1, C2, K, seed=s)[:2] for s in (1, 2, 3)]
    pinned = [(torch.from_numpy(a).pin_memory(), torch.from_numpy(b).pin_memory()) for a, b in scenes]
    want = [vitcnn_b200.predict_scene_host(net, a, b, device=DEV, chunk=900) for a, b in pinned]
    outs = [(torch.zeros(H, W, K).pin_memory(), torch.zeros(H, W, dtype=torch.uint8).pin_memory()) for _ in range(3)]
    for (a, b), (lg, am) in zip(pinned, outs):
        vitcnn_b200.predict_scene_host(net, a, b, device=DEV, chunk=900, logits_out=lg, argmax_out=am, sync=False)
    ev = vitcnn_b200.predict_scene_host.last_event(DEV)
    assert ev is not None
    ev.synchronize()
    torch.cuda.synchronize()
    for (lg, am), (wl, wa) in zip(outs, want):
        assert (wl != 0).any() and torch.equal(lg, wl) and torch.equal(am, wa)


@pytest.mark.parametrize("C,P", [(180, 11), (64, 7), (8, 5), (4, 9), (144, 15), (36, 11)])
def test_tma_gather_bit_exact_on_partial_channel_groups(C, P):
    """The tensor-map gather works in 32-channel groups (the last one zero-filled past C and not stored); P = 15 does not
    fit its stages and takes the generic kernel.  Plain and augmented samples against numpy indexing."""
    from vitcnn_b200 import ops
    H, W, n = 40, 53, 70
    rng = np.random.default_rng(C * 100 + P)
    img1 = rng.random((H, W, C), dtype=np.float32)
    img2 = rng.random((H, W, 2), dtype=np.float32)
    gt = rng.integers(0, 9, size=(H, W)).astype(np.uint8)
    p = P // 2
    xy = np.stack([rng.integers(p, H - P + p + 1, n), rng.integers(p, W - P + p + 1, n)], 1).astype(np.int32)
    codes = rng.integers(0, 7, n).astype(np.uint8)
    t1, t2, tg = (torch.from_numpy(a).to(DEV) for a in (img1, img2, gt))
    for use_ops in (False, True):
        hsi, lid, lab = ops.gather_patches(t1, t2, torch.from_numpy(xy).to(DEV), P, center_mode=True, gt=tg,
                                           ops=torch.from_numpy(codes).to(DEV) if use_ops else None)
        for b in range(n):
            x, y = int(xy[b, 0]) - p, int(xy[b, 1]) - p
            w1, w2 = img1[x:x + P, y:y + P], img2[x:x + P, y:y + P]
            if use_ops:
                w1, w2 = R.dihedral_apply(w1, int(codes[b])), R.dihedral_apply(w2, int(codes[b]))
            assert hsi[b].cpu().numpy().tobytes() == np.ascontiguousarray(w1.transpose(2, 0, 1)).tobytes(), (b, use_ops)
            assert lid[b].cpu().numpy().tobytes() == np.ascontiguousarray(w2.transpose(2, 0, 1)).tobytes(), (b, use_ops)


def test_gather_rejects_windows_that_leave_the_raster():
    from vitcnn_b200 import ops
    t1, t2 = torch.rand(20, 30, 8, device=DEV), torch.rand(20, 30, 1, device=DEV)
    with pytest.raises(ValueError):
        ops.gather_patches(t1, t2, torch.tensor([[4, 5]], dtype=torch.int32), 11)
    hsi, _, _ = ops.gather_patches(t1, t2, torch.tensor([[4, 5]], dtype=torch.int32), 11, validate=False)   # memory safe: moved inside
    assert torch.equal(hsi[0], t1[0:11, 0:11].permute(2, 0, 1))


_CHILD = r"""
import sys, torch
sys.path.insert(0, {root!r})
from oracle import data_ref as R
from tests.test_gpu_model import make_pair
H, W, C1, C2, P, K = 61, 75, 32, 1, {P}, 6
_, net = make_pair(C1, C2, P, K, seed=3)
img1, img2, _ = R.synthetic_scene(H, W, C1, C2, K, seed=4)
lg, am = net.predict_scene(torch.from_numpy(img1).cuda(), torch.from_numpy(img2).cuda(), chunk=1500)
torch.save((lg.cpu(), am.cpu()), sys.argv[1])
"""


def _scene_in_child(tmp_path, name, env, P=11):
    out = str(tmp_path / f"{name}.pt")
    e = dict(os.environ)
    e.update(env)
    res = subprocess.run([sys.executable, "-c", _CHILD.format(root=ROOT, P=P), out], env=e, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]
    return torch.load(out)


def test_kernel_variants_selected_by_environment(tmp_path):
    """Same scene through (a) the defaults, (b) the round-1 per-variant launches of the shared stem, (c) no LiDAR sharing,
    (d) sharing depth 2 with per-window conv 3 and the variant gather: all bit-identical maps.  (e) the other token kernels
    (three patches in flight, the shared-memory predecessor, two threads per row, mma.sync) sum in different orders: equal
    within the kernel tolerance."""
    base_l, base_a = _scene_in_child(tmp_path, "base", {})
    assert (base_l != 0).any()
    for name, env in (("planes", {"VITCNN_STEM_IMPL": "planes"}), ("nolidar", {"VITCNN_LIDAR_SHARED": "0"}),
                      ("depth2", {"VITCNN_SCENE_DEPTH": "2"}), ("depth0", {"VITCNN_SCENE_DEPTH": "0"})):
        lg, am = _scene_in_child(tmp_path, name, env)
        assert torch.equal(lg, base_l) and torch.equal(am, base_a), name
    scale = base_l.abs().max().item()
    # the token-kernel family: tokens_tm_kernel with three patches in flight, tokens_tc_kernel alone (the exact-softmax
    # fallback run unconditionally), its two-threads-per-row variant, the mma.sync kernel
    for name, env in (("tm3", {"VITCNN_TC_KERNEL": "tm3"}), ("tc", {"VITCNN_TC_KERNEL": "tc"}), ("split3", {"VITCNN_TC_SPLIT": "3"}),
                      ("split1", {"VITCNN_TC_SPLIT": "1"}), ("mma", {"VITCNN_TOKENS_IMPL": "0"})):
        lg, am = _scene_in_child(tmp_path, name, env)
        assert (lg - base_l).abs().max().item() <= 5e-3 * scale, name
        assert (lg == 0).eq(base_l == 0).all(), name

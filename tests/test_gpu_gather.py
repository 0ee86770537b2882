"""-m gpu: exact patch extraction through the C-ABI vs the numpy oracle and the committed
reference-generated fixtures.  Bit-exact (fp32 copies, int64 labels)."""
import numpy as np
import pytest
import torch

from oracle import data_ref as R

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _dev():
    return torch.device("cuda:0")


def _gather(img1, img2, gt, xy, P, center):
    from vitcnn_b200 import ops
    d = _dev()
    h, l, lab = ops.gather_patches(torch.from_numpy(img1).to(d), torch.from_numpy(img2).to(d),
                                   torch.from_numpy(np.ascontiguousarray(xy, dtype=np.int32)).to(d), P,
                                   center_mode=center, gt=None if gt is None else torch.from_numpy(gt).to(d))
    torch.cuda.synchronize()
    return h.cpu().numpy(), l.cpu().numpy(), None if lab is None else lab.cpu().numpy()


def test_golden_multimodalx_samples(data_golden):
    g = data_golden
    for ci, (H, W, C1, C2, P, n) in enumerate(g["ds_cases"]):
        idx = g[f"ds{ci}_indices"][:n]
        h, l, lab = _gather(g[f"ds{ci}_img1"], g[f"ds{ci}_img2"], g[f"ds{ci}_gt"], idx, int(P), True)
        assert h.tobytes() == g[f"ds{ci}_hsi"].tobytes()
        assert l.tobytes() == g[f"ds{ci}_lidar"].tobytes()
        assert np.array_equal(lab, g[f"ds{ci}_label"])


@pytest.mark.parametrize("P", [1, 5, 7, 8, 9, 11, 15])
@pytest.mark.parametrize("C1,C2", [(144, 1), (64, 2), (180, 1), (7, 3)])
def test_random_shapes_vs_oracle(P, C1, C2):
    rng = np.random.default_rng(P * 1000 + C1)
    H, W = P + 9, P + 14
    img1 = rng.random((H, W, C1), dtype=np.float32)
    img2 = rng.random((H, W, C2), dtype=np.float32)
    gt = rng.integers(0, 9, size=(H, W)).astype(np.int64)
    corners = R.sliding_window_corners((H, W), 1, (P, P))
    h, l, _ = _gather(img1, img2, None, corners, P, False)
    wh, wl = R.gather_corners(img1, img2, corners, P)
    assert h.tobytes() == wh.tobytes() and l.tobytes() == wl.tobytes()
    centers = corners + P // 2
    h, l, lab = _gather(img1, img2, gt, centers, P, True)
    assert h.tobytes() == wh.tobytes() and l.tobytes() == wl.tobytes()
    assert np.array_equal(lab, gt[centers[:, 0], centers[:, 1]])


def test_houston_size_checksum():
    """Full-size property: every window of a Houston-shaped row band, checked through sums
    the oracle can produce without materialising the patches (a patch sum is a box filter)."""
    from vitcnn_b200 import ops
    rng = np.random.default_rng(3)
    H, W, C1, C2, P = 40, 1905, 144, 1, 11
    img1 = rng.integers(0, 256, size=(H, W, C1)).astype(np.float32)   # integers: fp32 sums are exact
    img2 = rng.integers(0, 256, size=(H, W, C2)).astype(np.float32)
    corners = R.sliding_window_corners((H, W), 1, (P, P))
    d = _dev()
    t1, t2 = torch.from_numpy(img1).to(d), torch.from_numpy(img2).to(d)
    xy = torch.from_numpy(corners.astype(np.int32)).to(d)
    tot = np.zeros(len(corners))
    for s in range(0, len(corners), 8192):
        h, l, _ = ops.gather_patches(t1, t2, xy[s:s + 8192], P, center_mode=False)
        tot[s:s + 8192] = (h.double().sum((1, 2, 3)) + l.double().sum((1, 2, 3))).cpu().numpy()
    plane = img1.astype(np.float64).sum(2) + img2.astype(np.float64).sum(2)
    ii = np.pad(plane.cumsum(0).cumsum(1), ((1, 0), (1, 0)))
    x, y = corners[:, 0], corners[:, 1]
    want = ii[x + P, y + P] - ii[x, y + P] - ii[x + P, y] + ii[x, y]
    assert np.array_equal(tot, want)


def test_empty_batch():
    from vitcnn_b200 import ops
    d = _dev()
    h, l, _ = ops.gather_patches(torch.rand(12, 12, 4, device=d), torch.rand(12, 12, 1, device=d),
                                 torch.zeros(0, 2, dtype=torch.int32, device=d), 5)
    assert h.shape == (0, 4, 5, 5) and l.shape == (0, 1, 5, 5)


def test_flip_rotate_augmentation_matches_reference_golden(data_golden):
    """flip_augmentation on: the dataset mirror draws the reference's RNG decisions and the gather
    kernel applies them as an index remap -- bit-exact with what the reference's own
    MultiModalX.__getitem__ returned (tests/golden/make_golden.py), labels included."""
    import vitcnn_b200
    g = data_golden
    for ci, (H, W, C1, C2, P, n) in enumerate(g["ds_cases"]):
        img1, img2, gt = g[f"ds{ci}_img1"], g[f"ds{ci}_img2"], g[f"ds{ci}_gt"]
        hp = dict(dataset="synthetic", patch_size=int(P), ignored_labels=[0], flip_augmentation=True,
                  radiation_augmentation=False, mixture_augmentation=False, center_pixel=True,
                  supervision="full", applyPCA=False, device=DEV)
        np.random.seed(100 + ci)                     # seed_torch(100 + ci) of make_golden.py
        ds = vitcnn_b200.MultiModalX(img1, img2, gt, **hp)
        h, l, y = ds.batch(np.arange(int(n)))        # the draws of n consecutive __getitem__ calls
        assert h.cpu().numpy().tobytes() == g[f"ds{ci}_aug_hsi"].tobytes(), ci
        assert l.cpu().numpy().tobytes() == g[f"ds{ci}_aug_lidar"].tobytes(), ci
        assert np.array_equal(y.cpu().numpy(), g[f"ds{ci}_aug_label"]), ci


def test_all_dihedral_ops_and_radiation_noise_vs_oracle():
    from vitcnn_b200 import ops
    import vitcnn_b200
    for P in (5, 8, 11):       # even P: the label moves with the transform
        img1, img2, gt = R.synthetic_scene(30, 34, 12, 2, 6, seed=P)
        idx = R.train_indices(gt, [0], P)[:21]
        codes = np.arange(21, dtype=np.uint8) % 7
        h, l, y = ops.gather_patches(torch.from_numpy(img1).to(DEV), torch.from_numpy(img2).to(DEV),
                                     torch.from_numpy(idx.astype(np.int32)).to(DEV), P, center_mode=True,
                                     gt=torch.from_numpy(gt).to(DEV), ops=torch.from_numpy(codes).to(DEV))
        for i, (x, yy) in enumerate(idx):
            wh, wl, wy = R.augmented_sample(img1, img2, gt, int(x), int(yy), P, int(codes[i]))
            assert h[i].cpu().numpy().tobytes() == wh.tobytes() and l[i].cpu().numpy().tobytes() == wl.tobytes()
            assert int(y[i]) == int(wy)
    # radiation noise: same draws, float64 arithmetic like numpy (datasets.py:528-532)
    P = 7
    img1, img2, gt = R.synthetic_scene(30, 34, 12, 2, 6, seed=1)
    hp = dict(dataset="synthetic", patch_size=P, ignored_labels=[0], flip_augmentation=False,
              radiation_augmentation=True, mixture_augmentation=False, center_pixel=True,
              supervision="full", applyPCA=False, device=DEV)
    np.random.seed(5)
    ds = vitcnn_b200.MultiModalX(img1, img2, gt, **hp)
    state = np.random.get_state()
    h, l, y = ds.batch(np.arange(60))
    np.random.set_state(state)
    hit = 0
    for i in range(60):
        x, yy = ds.indices[i]
        wh, wl, wy = R.augmented_sample(img1, img2, gt, int(x), int(yy), P, 0)
        if np.random.random() < 0.1:
            alpha = np.random.uniform(0.9, 1.1)
            noise = np.random.normal(loc=0.0, scale=1.0, size=(P, P, 12))
            d = R.radiation_noise(img1[x - P // 2:x - P // 2 + P, yy - P // 2:yy - P // 2 + P], alpha, noise)
            wh = np.asarray(np.copy(d).transpose((2, 0, 1)), dtype="float32")
            hit += 1
        assert h[i].cpu().numpy().tobytes() == wh.tobytes(), i
    assert hit > 0


def test_minmax_normalise_bit_exact():
    """Raster ingest (datasets.py:124-133): per-band (HSI) and whole-array (LiDAR) min-max to [0,1]."""
    from vitcnn_b200.utils import minmax_normalise_
    rng = np.random.default_rng(4)
    for H, W, C in ((37, 53, 144), (20, 31, 1), (19, 23, 5)):
        raw = (rng.random((H, W, C), dtype=np.float32) * 4000 - 300).astype(np.float32)
        want = R.minmax_normalise(raw)
        got = minmax_normalise_(torch.from_numpy(raw.copy()).to(DEV)).cpu().numpy()
        assert got.tobytes() == want.tobytes()
        lo, hi = raw.min(), raw.max()
        want_g = (raw - lo) / (hi - lo)
        got_g = minmax_normalise_(torch.from_numpy(raw.copy()).to(DEV), per_band=False).cpu().numpy()
        assert got_g.tobytes() == want_g.tobytes()

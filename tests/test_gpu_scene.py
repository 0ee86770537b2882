"""-m gpu: full-scene sliding-window inference (the loop of test(), model_utils.py:1067-1132)
vs the oracle's restatement of that loop driving the fp32 oracle model."""
import numpy as np
import pytest
import torch

from oracle import data_ref as R
from tests.test_gpu_model import DEV, REL_TOL, make_pair

pytestmark = pytest.mark.gpu


def _oracle_probs(ref, img1, img2, P, K, stride, bs=64):
    def net(h, l):
        with torch.no_grad():
            return ref(torch.from_numpy(h), torch.from_numpy(l)).numpy()
    return R.scene_test(net, img1, img2, P, bs, K, stride)


@pytest.mark.parametrize("cfg", [(24, 31, 16, 1, 5, 4, 1), (30, 26, 64, 2, 7, 12, 2), (27, 40, 144, 1, 11, 16, 1),
                                 (29, 33, 20, 1, 9, 5, 3)])
def test_scene_vs_oracle_loop(cfg):
    import vitcnn_b200
    H, W, C1, C2, P, K, stride = cfg
    ref, ours = make_pair(C1, C2, P, K)
    img1, img2, _ = R.synthetic_scene(H, W, C1, C2, K, seed=1)
    want = _oracle_probs(ref, img1, img2, P, K, stride)
    hp = dict(patch_size=P, center_pixel=True, batch_size=64, device=torch.device(DEV), n_classes=K,
              applyPCA=False, test_stride=stride, scene_chunk=100)   # several chunks, short last one
    got = vitcnn_b200.test(0, ours, img1, img2, hp)
    assert got.dtype == np.float64 and got.shape == want.shape
    assert np.array_equal(got == 0, want == 0)          # untouched pixels stay exactly zero
    assert np.abs(got - want).max() <= REL_TOL * np.abs(want).max()


def test_scene_equals_batched_forward_and_row_bands():
    """The scene path must give bit-identical logits to forward() on the same windows, and a
    row-band split (any number of ranks) must reproduce the single-pass map bit for bit."""
    from vitcnn_b200.utils import row_band_ranges, window_starts
    H, W, C1, C2, P, K = 33, 45, 144, 1, 11, 16
    ref, ours = make_pair(C1, C2, P, K)
    img1, img2, _ = R.synthetic_scene(H, W, C1, C2, K, seed=4)
    t1, t2 = torch.from_numpy(img1).to(DEV), torch.from_numpy(img2).to(DEV)
    full, amax = ours.predict_scene(t1, t2, chunk=257)
    corners = R.sliding_window_corners((H, W), 1, (P, P))
    h, l = R.gather_corners(img1, img2, corners, P)
    with torch.no_grad():
        lg = ours(torch.from_numpy(h).to(DEV), torch.from_numpy(l).to(DEV))
    cx, cy = corners[:, 0] + P // 2, corners[:, 1] + P // 2
    assert torch.equal(full[cx, cy], lg)
    assert torch.equal(amax[cx, cy].long(), lg.argmax(1))
    nx, ny = len(window_starts(H, P, 1)), len(window_starts(W, P, 1))
    for world in (2, 3, 8):
        acc = torch.zeros_like(full)
        am = torch.zeros_like(amax)
        for first, count in row_band_ranges(nx, ny, world):
            ours.predict_scene(t1, t2, chunk=300, window_range=(first, count), logits_map=acc, argmax_map=am)
        assert torch.equal(acc, full) and torch.equal(am, amax)


def test_device_metrics_match_reference_formulas():
    """metrics() (utils.py:585-663) with the confusion matrix counted on the device: integers
    bit-exact, derived floats equal to the oracle's restatement (pinned to the reference)."""
    from vitcnn_b200.utils import metrics
    rng = np.random.default_rng(3)
    H, W, K = 61, 83, 16
    gt = rng.integers(0, K, size=(H, W)).astype(np.int64)
    pred = np.where(rng.random((H, W)) < 0.7, gt, rng.integers(0, K, size=(H, W))).astype(np.int64)
    for ignored, n_classes in (([0], K), ([0, 3], K), ([], None)):
        want = R.metrics(pred, gt, ignored_labels=ignored, n_classes=n_classes)
        got = metrics(torch.from_numpy(pred.astype(np.uint8)).to(DEV), torch.from_numpy(gt).to(DEV),
                      ignored_labels=ignored, n_classes=n_classes)
        assert np.array_equal(got["Confusion matrix"], want["Confusion matrix"])
        for k in ("Accuracy", "AA", "Kappa"):
            assert got[k] == want[k], k
        for k in ("F1 scores", "Precisions"):
            assert np.array_equal(got[k], want[k], equal_nan=True), k


@pytest.mark.parametrize("cfg", [(61, 47, 144, 1, 11, 16), (40, 95, 64, 2, 8, 12), (15, 15, 20, 1, 2, 5), (44, 29, 180, 1, 15, 8),
                                 (28, 41, 16, 1, 3, 4), (100, 100, 16, 1, 7, 4), (57, 120, 32, 1, 9, 6), (90, 40, 16, 2, 15, 3)])
def test_shared_stem_is_bit_identical_to_the_per_window_path(cfg):
    """Dense scenes compute the border-class variants of the HSI stem convs once per scene (all three convs on
    31 x 31 blocks when P >= 7 and the raster allows, conv 1 only on 15 x 15 blocks otherwise; blocks clamped at
    the far edges) and gather every window's stem output from them; the logits map must equal the per-window
    path bit for bit (the per-window path = the same library call without w_h1_border)."""
    H, W, C1, C2, P, K = cfg
    _, ours = make_pair(C1, C2, P, K, seed=H)
    img1, img2, _ = R.synthetic_scene(H, W, C1, C2, K, seed=W)
    t1, t2 = torch.from_numpy(img1).to(DEV), torch.from_numpy(img2).to(DEV)
    shared, am_s = ours.predict_scene(t1, t2, chunk=500)
    st = ours.pack_for_inference()["struct"]
    keep = st.w_h1_border
    assert keep, "the model must ship the border-class weight copies"
    st.w_h1_border = None
    try:
        plain, am_p = ours.predict_scene(t1, t2, chunk=500)
    finally:
        st.w_h1_border = keep
    torch.cuda.synchronize()
    assert (shared != 0).any()
    assert torch.equal(shared, plain) and torch.equal(am_s, am_p)


@pytest.mark.parametrize("split", ["blocks", "equal"])
def test_host_pipeline_sub_bands_reproduce_the_single_pass_map(split, monkeypatch):
    """predict_scene_host cuts the band into sub-bands (whole block rows of the shared stem, or equal parts) and
    pipelines upload / compute / download; whatever the cut, the maps must equal predict_scene's bit for bit."""
    import vitcnn_b200
    monkeypatch.setenv("VITCNN_SUBBAND_SPLIT", split)
    H, W, C1, C2, P, K = 131, 52, 32, 1, 11, 6
    _, ours = make_pair(C1, C2, P, K, seed=7)
    img1, img2, _ = R.synthetic_scene(H, W, C1, C2, K, seed=9)
    t1, t2 = torch.from_numpy(img1), torch.from_numpy(img2)
    full, amax = ours.predict_scene(t1.to(DEV), t2.to(DEV), chunk=700)
    for world in (1, 2):
        lg = torch.zeros(H, W, K).pin_memory()
        am = torch.zeros(H, W, dtype=torch.uint8).pin_memory()
        for rank in range(world):
            vitcnn_b200.predict_scene_host(ours, t1.pin_memory(), t2.pin_memory(), rank=rank, world=world, chunk=700,
                                           logits_out=lg, argmax_out=am, device=DEV)
        assert torch.equal(lg, full.cpu()) and torch.equal(am, amax.cpu())

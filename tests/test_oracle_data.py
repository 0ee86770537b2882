"""Pin the numpy oracle (oracle/data_ref.py) to fixtures produced by the reference's
own utils.sliding_window / datasets.MultiModalX / model_utils.test / val / utils.metrics
(tests/golden/make_golden.py).  Bit-exact: these are index / copy / integer paths."""
import numpy as np
import pytest

from oracle import data_ref as R
from oracle import ref_import


def test_sliding_window_corners(data_golden):
    g = data_golden
    for H, W, P, step in g["sw_cases"]:
        want = g[f"sw_{H}_{W}_{P}_{step}"]
        got = R.sliding_window_corners((H, W), int(step), (int(P), int(P)))
        assert got.shape == want.shape and np.array_equal(got, want), (H, W, P, step)
        assert R.count_sliding_window((H, W), int(step), (int(P), int(P))) == len(want)


def test_grouper(data_golden):
    assert [len(c) for c in R.grouper(7, range(23))] == list(data_golden["grouper_7_of_23"])
    assert list(R.grouper(4, [])) == []


def test_multimodalx_indices_and_samples(data_golden):
    g = data_golden
    for ci, (H, W, C1, C2, P, n) in enumerate(g["ds_cases"]):
        img1, img2, gt = g[f"ds{ci}_img1"], g[f"ds{ci}_img2"], g[f"ds{ci}_gt"]
        np.random.seed(ci)          # seed_torch(ci) -> np.random.seed(ci), utils.py:890
        idx = R.shuffled_train_indices(gt, [0], int(P))
        assert np.array_equal(idx, g[f"ds{ci}_indices"])
        hsi, lid, lab = R.gather_centers(img1, img2, gt, idx[:n], int(P))
        assert hsi.tobytes() == g[f"ds{ci}_hsi"].tobytes()
        assert lid.tobytes() == g[f"ds{ci}_lidar"].tobytes()
        assert np.array_equal(lab, g[f"ds{ci}_label"])


def test_strict_border_rule():
    # datasets.py:497-504: p < x < H-p, strict on both sides (row p is excluded)
    gt = np.ones((12, 13), np.uint8)
    idx = R.train_indices(gt, [0], 5)
    assert idx[:, 0].min() == 3 and idx[:, 0].max() == 12 - 2 - 1
    assert idx[:, 1].min() == 3 and idx[:, 1].max() == 13 - 2 - 1


def test_flip_rotate_augmentation(data_golden):
    g = data_golden
    for ci, (H, W, C1, C2, P, n) in enumerate(g["ds_cases"]):
        img1, img2, gt = g[f"ds{ci}_img1"], g[f"ds{ci}_img2"], g[f"ds{ci}_gt"]
        np.random.seed(100 + ci)
        idx = R.shuffled_train_indices(gt, [0], int(P))
        for i in range(n):
            op = R.draw_spatial_aug_op()
            hsi, lid, lab = R.augmented_sample(img1, img2, gt, int(idx[i, 0]), int(idx[i, 1]), int(P), op)
            assert hsi.tobytes() == g[f"ds{ci}_aug_hsi"][i].tobytes(), (ci, i, op)
            assert lid.tobytes() == g[f"ds{ci}_aug_lidar"][i].tobytes()
            assert lab == g[f"ds{ci}_aug_label"][i]


def _toy(K):
    from tests.golden.make_golden import toy_net_numpy
    return lambda h, l: toy_net_numpy(h, l, K)


def test_scene_test_scatter(data_golden):
    g = data_golden
    for ti, (H, W, C1, C2, P, stride, bs) in enumerate(g["test_cases"]):
        probs = R.scene_test(_toy(5), g[f"test{ti}_img1"], g[f"test{ti}_img2"], int(P), int(bs), 5, int(stride))
        want = g[f"test{ti}_probs"]
        assert probs.dtype == np.float64 and probs.shape == want.shape
        # toy net is fp32 arithmetic in torch vs numpy: same values to fp32 rounding
        np.testing.assert_allclose(probs, want, rtol=2e-6, atol=2e-6)
        assert np.array_equal(probs == 0, want == 0)      # untouched border stays exactly 0


def test_val_accuracy(data_golden):
    g = data_golden
    img1, gt, idx = g["val_img1"], g["val_gt"], g["val_indices"]
    P, K = 5, 5
    pred = (img1[idx[:, 0], idx[:, 1], 0] * np.float32(1000)).astype(np.int64) % K
    tgt = gt[idx[:, 0], idx[:, 1]]
    assert R.val_accuracy(pred, tgt, {0}) == pytest.approx(float(g["val_acc"]), abs=0)


def test_metrics(data_golden):
    g = data_golden
    with np.errstate(all="ignore"):
        res = R.metrics(g["met_pred"], g["met_tgt"], ignored_labels=[0], n_classes=5)
    assert np.array_equal(res["Confusion matrix"], g["met_cm"])
    assert res["Accuracy"] == pytest.approx(float(g["met_acc"]), rel=1e-12)
    np.testing.assert_allclose(res["F1 scores"], g["met_f1"], rtol=1e-12, equal_nan=True)
    np.testing.assert_allclose(res["Precisions"], g["met_prec"], rtol=1e-12, equal_nan=True)
    assert res["AA"] == pytest.approx(float(g["met_aa"]), rel=1e-12)
    assert res["Kappa"] == pytest.approx(float(g["met_kappa"]), rel=1e-12)


@pytest.mark.skipif(not ref_import.available(), reason="reference tree not mounted")
def test_live_against_reference_random_shapes():
    """Dev-container only: hypothesis-style sweep against the live reference."""
    ref_utils, ref_datasets, _ = ref_import.import_reference(with_model_utils=False)
    rng = np.random.default_rng(0)
    for _ in range(25):
        P = int(rng.choice([1, 3, 5, 7, 8, 9, 11, 15]))
        H, W = int(rng.integers(P, P + 20)), int(rng.integers(P, P + 20))
        step = int(rng.integers(1, 4))
        a = np.zeros((H, W, 1), np.float32)
        want = np.array([(x, y) for x, y, w, h in
                         ref_utils.sliding_window(a, a, step=step, window_size=(P, P), with_data=False)],
                        dtype=np.int64).reshape(-1, 2)
        assert np.array_equal(R.sliding_window_corners((H, W), step, (P, P)), want)

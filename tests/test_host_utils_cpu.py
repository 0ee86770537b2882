"""CPU tests of the host-side mirrors around the hot path: sample_gt against splits the
reference's own sample_gt produced (tests/golden/split_golden.npz, made by make_split_golden.py) and the checkpoint
writer's naming / round trip."""
import os

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(HERE, "golden", "split_golden.npz"))


@pytest.mark.parametrize("tag,mode,ts", [("random_frac", "random", 0.2), ("random_count", "random", 50),
                                         ("disjoint_50", "disjoint", 0.5), ("disjoint_30", "disjoint", 0.3)])
def test_sample_gt_matches_reference_splits(golden, tag, mode, ts):
    from vitcnn_b200.utils import sample_gt
    np.random.seed(7)
    tr, te = sample_gt(golden["gt"], ts, mode=mode)
    assert tr.dtype == golden[tag + "_train"].dtype
    assert np.array_equal(tr, golden[tag + "_train"]) and np.array_equal(te, golden[tag + "_test"])


def test_fixed_number_split_matches_reference_draw(golden):
    from vitcnn_b200.utils import _fixed_number_split, sample_gt
    gt = golden["gt"]
    tr, te = _fixed_number_split(5, gt.reshape(-1).astype(np.int64), 11)
    assert np.array_equal(tr, golden["fixednum_train_idx"]) and np.array_equal(te, golden["fixednum_test_idx"])
    a, b = sample_gt(gt, 5, mode="random_fixednumber", seed=11)       # upstream dies on np.int here (utils.py:836)
    assert a.shape == gt.shape and a.dtype == np.float64
    assert all(int((a == c).sum()) == 5 for c in range(1, int(gt.max()) + 1))
    assert not ((a > 0) & (b > 0)).any() and np.array_equal(a + b, gt.astype(np.float64))


def test_sample_gt_fixed_mode_and_errors(golden):
    from vitcnn_b200.utils import sample_gt
    gt = golden["gt"]
    np.random.seed(2)
    tr, te = sample_gt(gt, 0.25, mode="fixed")
    assert np.array_equal(tr + te, gt) and not ((tr > 0) & (te > 0)).any()
    for c in range(1, int(gt.max()) + 1):
        n = int((gt == c).sum())
        assert int((tr == c).sum()) == int(0.25 * n)           # sklearn floors the train share
    with pytest.raises(ValueError):
        sample_gt(gt, 0.5, mode="nope")
    with pytest.raises(ValueError):
        sample_gt(gt.reshape(-1), 0.5)


def test_save_model_naming_and_round_trip(tmp_path, monkeypatch):
    import vitcnn_b200
    from vitcnn_b200.model_utils import save_model
    from oracle.model_ref import ViTCNNRef
    monkeypatch.chdir(tmp_path)
    torch.manual_seed(0)
    net = vitcnn_b200.ViTCNN(12, 1, patch_size=5, num_classes=4)
    path = save_model("ViTCNN", net, "ViTCNN", "Houston2013", "train", "best", run=2, epoch=7, metric=91.256)
    assert path.startswith("./checkpoints/ViTCNN/Houston2013/train/best/") and path.endswith("ViTCNN_run2_epoch7_91.26.pth")
    ref = ViTCNNRef(12, 1, patch_size=5, num_classes=4)
    ref.load_state_dict(torch.load(path))                        # the reference-style module takes our checkpoint as it is
    sd = net.state_dict()
    assert all(torch.equal(v, sd[k]) for k, v in ref.state_dict().items())
    other = save_model("x", {"a": 1}, "SVM", "d", "train", "last")
    assert other.endswith(".pkl") and os.path.isfile(other)


@pytest.mark.parametrize("P", [7, 9, 11, 15])
@pytest.mark.parametrize("block,depth", [(31, 2), (31, 3), (15, 1), (0, 0)])
def test_subband_bounds_cover_the_band_in_whole_block_rows(P, block, depth):
    """Sub-bands of the host pipeline (scene.py): contiguous, cover every window row once, at most `pipeline`
    of them, and cut so that the shared stem's block rows (csrc/vc_common.cuh blk_count) never exceed what
    equal parts would need."""
    from vitcnn_b200.utils import subband_bounds

    def block_rows(wr):          # blk_count of a sub-band of wr window rows (raster rows = wr + P - 1)
        ext = wr + P - 1
        return 1 if ext <= block else (ext - block + block - 2 * depth - 1) // (block - 2 * depth) + 1

    for nrows in list(range(1, 130)) + [339, 340, 1000]:
        for pipeline in (1, 2, 6, 8):
            b = subband_bounds(nrows, P, pipeline, block, depth)
            assert b[0] == 0 and b[-1] == nrows and len(b) - 1 <= pipeline
            sizes = np.diff(b)
            assert (sizes > 0).all()
            if block and block - (P - 1) > 0:
                nsub = max(1, min(pipeline, nrows // 21))
                equal = [(nrows * k) // nsub for k in range(nsub + 1)]
                assert sum(block_rows(s) for s in sizes) <= sum(block_rows(s) for s in np.diff(equal)) + (len(sizes) > nsub)
    assert subband_bounds(339, 11, 8, 31, 2) == [0, 21, 69, 117, 165, 213, 261, 309, 339]      # Houston: 15 block rows
    assert subband_bounds(339, 11, 6, 31, 2) == [0, 48, 96, 144, 192, 267, 339]                # 14 block rows
    assert subband_bounds(339, 11, 6, 31, 2, lead_small=True) == [0, 21, 69, 117, 192, 267, 339]


def test_plan_subbands_picks_whole_block_rows_and_covers_every_window_row():
    """utils.plan_subbands (scene.py): the spans cover every window row, only ever overlap by starting earlier, never
    exceed the allowed count, and for the Houston scene the single-scene plan is four sub-bands of exactly one 95-row
    block row each (the uncut scene needs four) while the streaming plan does not cut at all.  The block edge is
    modelled as csrc/abi.cu scene_block does it (fewest block-row pixels among 31 / 63 / 95)."""
    from vitcnn_b200.utils import plan_subbands
    W, D = 1905, 3

    def blk_count(ext, B):
        st = B - 2 * D
        return (ext - B + st - 1) // st + 1

    def block_of(h, depth):
        best, rows = 31, -1
        for B in (31, 63, 95):
            if h < B or W < B:
                break
            r = blk_count(h, B) * blk_count(W, B) * (B + 1) * (B + 1)
            if rows < 0 or r < rows:
                best, rows = B, r
        return best

    P = 11
    cost, spans = plan_subbands(339, P, 6, True, lambda rows: D, block_of)
    assert spans == [(0, 85), (85, 170), (170, 255), (254, 339)]
    assert all(block_of(b - a + P - 1, D) == 95 and blk_count(b - a + P - 1, 95) == 1 for a, b in spans)
    assert plan_subbands(339, P, 2, False, lambda rows: D, block_of)[1] == [(0, 339)]
    assert plan_subbands(339, P, 6, True, lambda rows: 0, block_of) is None          # no shared stem: the caller keeps equal parts
    for nrows in list(range(21, 200, 7)) + [339, 340, 700]:
        for nmax in (1, 2, 4, 6):
            for sync in (True, False):
                plan = plan_subbands(nrows, P, min(nmax, max(1, nrows // 21)), sync, lambda rows: D, block_of)
                assert plan is not None
                spans = plan[1]
                assert 1 <= len(spans) <= nmax and spans[0][0] == 0 and spans[-1][1] == nrows
                covered = 0
                for a, b in spans:
                    assert 0 <= a < b <= nrows and a <= covered          # contiguous or overlapping, never a gap
                    covered = max(covered, b)
                assert covered == nrows

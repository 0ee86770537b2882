"""CPU oracle for the data side of the ViT-CNN hot path (numpy restatement).

TEST INFRASTRUCTURE ONLY.  Nothing under ``vit-cnn_b200/`` may import this
module; it is imported by ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` as the checker.

Every function restates one piece of the reference toolkit (paths relative to
``/root/reference``) and is pinned bit-exactly against the reference's own
functions by ``tests/golden/make_golden.py`` -> ``tests/golden/*.npz`` ->
``tests/test_oracle_data.py``.

Raster convention: ``img1`` f32 [H, W, C1], ``img2`` f32
[H, W, C2], ``gt`` int [H, W]; axis 0 is called ``x`` / ``W`` by the reference
although it is the row axis.
"""
from __future__ import annotations

import itertools

import numpy as np


# --------------------------------------------------------------------------
# Sliding window enumeration (utils.py:357-415, 567-582)
# --------------------------------------------------------------------------
def window_starts(extent: int, win: int, step: int) -> np.ndarray:
    """Start offsets along one axis, utils.py:374-397.

    ``range(0, extent - win + offset + 1, step)`` with ``offset = (extent - win)
    % step``; a start that overshoots is clamped to ``extent - win`` so the last
    window is always included (and may repeat the previous one's pixels).
    """
    offset = (extent - win) % step
    starts = np.arange(0, extent - win + offset + 1, step, dtype=np.int64)
    starts[starts + win > extent] = extent - win
    return starts


def sliding_window_corners(shape_hw, step: int, window_size) -> np.ndarray:
    """All (x, y) top-left corners in the reference's generator order.

    utils.py:385-397: outer loop over axis 0 (``x``), inner over axis 1 (``y``).
    Returns int64 [N, 2].
    """
    w, h = window_size
    xs = window_starts(shape_hw[0], w, step)
    ys = window_starts(shape_hw[1], h, step)
    if xs.size == 0 or ys.size == 0:
        return np.zeros((0, 2), dtype=np.int64)
    xx, yy = np.meshgrid(xs, ys, indexing="ij")
    return np.stack([xx.ravel(), yy.ravel()], axis=1)


def count_sliding_window(shape_hw, step: int, window_size) -> int:
    """utils.py:404-415."""
    return int(sliding_window_corners(shape_hw, step, window_size).shape[0])


def grouper(n: int, iterable):
    """utils.py:567-582 - consecutive chunks of at most ``n`` items."""
    it = iter(iterable)
    while True:
        chunk = tuple(itertools.islice(it, n))
        if not chunk:
            return
        yield chunk


# --------------------------------------------------------------------------
# Training sample set (datasets.py:464-508)
# --------------------------------------------------------------------------
def train_indices(gt: np.ndarray, ignored_labels, patch_size: int,
                  supervision: str = "full") -> np.ndarray:
    """Un-shuffled labelled-pixel list of MultiModalX.__init__ (datasets.py:489-504).

    mask = gt not in ignored; ``np.nonzero`` row-major; keep (x, y) with
    ``p < x < H - p`` and ``p < y < W - p`` (strict on both sides).
    """
    if supervision == "full":
        mask = np.ones_like(gt)
        for l in set(ignored_labels):
            mask[gt == l] = 0
    elif supervision == "semi":
        mask = np.ones_like(gt)
    else:
        raise ValueError(supervision)
    x_pos, y_pos = np.nonzero(mask)
    p = patch_size // 2
    keep = (x_pos > p) & (x_pos < gt.shape[0] - p) & (y_pos > p) & (y_pos < gt.shape[1] - p)
    return np.stack([x_pos[keep], y_pos[keep]], axis=1).astype(np.int64)


def shuffled_train_indices(gt, ignored_labels, patch_size, supervision="full"):
    """datasets.py:506 - ``np.random.shuffle`` on the [N, 2] array under the
    caller's numpy seed (``seed_torch``, utils.py:887-895).

    The reference first builds ``self.labels`` with one ``self.label[x, y]`` read
    per index (no RNG use), so the RNG stream position is unchanged by it.
    """
    idx = train_indices(gt, ignored_labels, patch_size, supervision)
    np.random.shuffle(idx)
    return idx


# --------------------------------------------------------------------------
# Patch extraction (datasets.py:550-593, model_utils.py:1103-1112)
# --------------------------------------------------------------------------
def extract_patch_center(img: np.ndarray, x: int, y: int, P: int) -> np.ndarray:
    """[C, P, P] f32 patch around centre (x, y), datasets.py:551-556,571."""
    x1, y1 = x - P // 2, y - P // 2
    return np.asarray(np.copy(img[x1:x1 + P, y1:y1 + P]).transpose((2, 0, 1)), dtype="float32")


def gather_centers(img1, img2, gt, centers: np.ndarray, P: int):
    """Batch of MultiModalX.__getitem__ results, stacked like default_collate.

    Returns (hsi [B, C1, P, P] f32, lidar [B, C2, P, P] f32, label [B] int64),
    label = gt[x, y] (center_pixel=True, datasets.py:580-581).
    """
    B = len(centers)
    hsi = np.empty((B, img1.shape[2], P, P), np.float32)
    lid = np.empty((B, img2.shape[2], P, P), np.float32)
    lab = np.empty((B,), np.int64)
    for b, (x, y) in enumerate(centers):
        hsi[b] = extract_patch_center(img1, int(x), int(y), P)
        lid[b] = extract_patch_center(img2, int(x), int(y), P)
        lab[b] = gt[x, y] if gt is not None else 0
    return hsi, lid, lab


def gather_corners(img1, img2, corners: np.ndarray, P: int):
    """test()'s batch assembly (model_utils.py:1103-1112): windows by top-left
    corner, values in NCHW index order (strides differ from the reference's
    NHWC-memory view, values do not - SURVEY.md App. C rule 5)."""
    B = len(corners)
    hsi = np.empty((B, img1.shape[2], P, P), np.float32)
    lid = np.empty((B, img2.shape[2], P, P), np.float32)
    for b, (x, y) in enumerate(corners):
        hsi[b] = img1[x:x + P, y:y + P].transpose(2, 0, 1)
        lid[b] = img2[x:x + P, y:y + P].transpose(2, 0, 1)
    return hsi, lid


# --------------------------------------------------------------------------
# Augmentations (datasets.py:510-532, 559-566)  -- row (f)2
# --------------------------------------------------------------------------
def dihedral_apply(arr: np.ndarray, op: int) -> np.ndarray:
    """Apply one of the reference's spatial augmentations to an [h, w, ...] array.

    op 0 identity; 1 fliplr; 2 flipud; 3 fliplr then flipud (datasets.py:511-518);
    4/5/6 rot90 with k = 1/2/3 (datasets.py:521-526).
    """
    if op == 0:
        return arr
    if op == 1:
        return np.fliplr(arr)
    if op == 2:
        return np.flipud(arr)
    if op == 3:
        return np.flipud(np.fliplr(arr))
    if op in (4, 5, 6):
        return np.rot90(arr, k=op - 3)
    raise ValueError(op)


def draw_spatial_aug_op() -> int:
    """The RNG draws of one ``__getitem__`` with flip_augmentation on
    (datasets.py:559-564 -> 510-526), from numpy's global RNG, as an op code
    for :func:`dihedral_apply`."""
    if np.random.random() > 0.5:
        horizontal = np.random.random() > 0.5
        vertical = np.random.random() > 0.5
        return int(horizontal) + 2 * int(vertical)
    if np.random.random() > 0.5:
        return 3 + int(np.random.choice([1, 2, 3]))
    return 0


def augmented_sample(img1, img2, gt, x, y, P, op):
    """One MultiModalX sample with spatial augmentation ``op`` applied to the data,
    LiDAR and label windows alike; the label is read at [P//2, P//2] AFTER the
    transform (datasets.py:557-581), which matters for even P."""
    x1, y1 = x - P // 2, y - P // 2
    d = dihedral_apply(img1[x1:x1 + P, y1:y1 + P], op)
    d2 = dihedral_apply(img2[x1:x1 + P, y1:y1 + P], op)
    lab = dihedral_apply(gt[x1:x1 + P, y1:y1 + P], op)
    hsi = np.asarray(np.copy(d).transpose((2, 0, 1)), dtype="float32")
    lid = np.asarray(np.copy(d2).transpose((2, 0, 1)), dtype="float32")
    return hsi, lid, np.int64(lab[P // 2, P // 2])


def radiation_noise(data, alpha, noise, beta=1 / 25):
    """datasets.py:528-532 with the RNG draws passed in (host draws them)."""
    return alpha * data + beta * noise


# --------------------------------------------------------------------------
# Scene inference (model_utils.py:1067-1132)
# --------------------------------------------------------------------------
def scene_test(net_fn, img1, img2, patch_size: int, batch_size: int, n_classes: int,
               test_stride: int = 1) -> np.ndarray:
    """Restates test() for center_pixel=True, patch_size > 1.

    ``net_fn(hsi[B,C1,P,P] f32, lidar[B,C2,P,P] f32) -> [B, K]`` array.
    probs is float64 [H, W, K]; window (x, y) adds its raw logits at
    ``probs[x + P//2, y + P//2]`` (model_utils.py:1127-1129); untouched pixels stay 0.
    """
    P = patch_size
    probs = np.zeros(img1.shape[:2] + (n_classes,))
    corners = sliding_window_corners(img1.shape[:2], test_stride, (P, P))
    for s in range(0, len(corners), batch_size):
        chunk = corners[s:s + batch_size]
        hsi, lid = gather_corners(img1, img2, chunk, P)
        out = np.asarray(net_fn(hsi, lid))
        for (x, y), o in zip(chunk, out):
            probs[x + P // 2, y + P // 2] += o
    return probs


# --------------------------------------------------------------------------
# Validation / metrics (model_utils.py:1135-1158, utils.py:585-663)
# --------------------------------------------------------------------------
def val_accuracy(pred: np.ndarray, target: np.ndarray, ignored_labels) -> float:
    """val(): predictions that fall in ``ignored_labels`` are skipped, the rest
    are scored against the target (model_utils.py:1152-1157)."""
    pred = np.asarray(pred).ravel()
    target = np.asarray(target).ravel()
    keep = ~np.isin(pred, list(ignored_labels))
    return float(np.sum(pred[keep] == target[keep])) / float(np.sum(keep))


def confusion_matrix(target: np.ndarray, prediction: np.ndarray, n_classes: int) -> np.ndarray:
    """sklearn.metrics.confusion_matrix(target, prediction, labels=range(n)):
    cm[t, p] counts; pairs with either label outside range are dropped."""
    target = np.asarray(target).ravel().astype(np.int64)
    prediction = np.asarray(prediction).ravel().astype(np.int64)
    ok = (target >= 0) & (target < n_classes) & (prediction >= 0) & (prediction < n_classes)
    cm = np.bincount(target[ok] * n_classes + prediction[ok], minlength=n_classes * n_classes)
    return cm.reshape(n_classes, n_classes).astype(np.int64)


def metrics(prediction, target, ignored_labels=(), n_classes=None) -> dict:
    """utils.py:585-663 (numpy warnings instead of ZeroDivisionError, as there)."""
    ignored = np.zeros(target.shape[:2], dtype=bool)
    for l in ignored_labels:
        ignored[target == l] = True
    keep = ~ignored
    t = target[keep]
    p = prediction[keep]
    n_classes = int(np.max(t)) + 1 if n_classes is None else n_classes
    cm = confusion_matrix(t, p, n_classes)
    total = np.sum(cm)
    res = {"Confusion matrix": cm}
    res["Accuracy"] = sum(cm[i][i] for i in range(len(cm))) * (100 / float(total))
    with np.errstate(divide="ignore", invalid="ignore"):
        rs, cs, dg = cm.sum(1), cm.sum(0), np.diag(cm)
        res["F1 scores"] = 2.0 * dg / (rs + cs)
        res["Precisions"] = 1.0 * dg / rs
        rec = dg / rs
        res["AA"] = np.mean(rec[~np.isnan(rec)])
        pa = np.trace(cm) / float(total)
        pe = np.sum(cs * rs) / float(total * total)
        res["Kappa"] = (pa - pe) / (1 - pe)
    return res


# --------------------------------------------------------------------------
# Raster normalisation (datasets.py:124-133)  -- row (f)3
# --------------------------------------------------------------------------
def minmax_normalise(img: np.ndarray) -> np.ndarray:
    """Per-band min-max to [0, 1] as float32 (datasets.py:124-133, 283-295, 321-332):
    computed band by band in the raster's own dtype, then cast."""
    img = np.asarray(img, dtype="float32")
    out = np.empty_like(img)
    for b in range(img.shape[2]):
        band = img[:, :, b]
        lo, hi = np.min(band), np.max(band)
        out[:, :, b] = (band - lo) / (hi - lo)
    return out


# --------------------------------------------------------------------------
# Synthetic inputs
# --------------------------------------------------------------------------
def synthetic_scene(H, W, C1, C2, K, seed=0, structured=True, block=8):
    """Synthetic co-registered rasters.

    structured=False: U[0,1) noise (throughput rasters).
    structured=True : blocky label map (labels 1..K-1, ~1/3 unlabelled 0), a smooth
    class spectrum plus noise, per-class LiDAR height plus noise, min-max per band.
    Returns (img1 f32 [H,W,C1], img2 f32 [H,W,C2], gt uint8 [H,W]).
    """
    rng = np.random.default_rng(seed)
    if not structured:
        img1 = rng.random((H, W, C1), dtype=np.float32)
        img2 = rng.random((H, W, C2), dtype=np.float32)
        gt = rng.integers(0, K, size=(H, W)).astype(np.uint8)
        return img1, img2, gt
    bs = int(block)      # edge of the label blocks (8 by default: every P >= 9 window straddles several classes)
    gh, gw = (H + bs - 1) // bs, (W + bs - 1) // bs
    blocks = rng.integers(1, K, size=(gh, gw))
    blocks[rng.random((gh, gw)) < 1 / 3] = 0
    gt = np.kron(blocks, np.ones((bs, bs), dtype=np.int64))[:H, :W].astype(np.uint8)
    spec = rng.random((K, C1))
    kern = np.ones(9) / 9.0
    spec = np.stack([np.convolve(np.pad(s, 4, mode="edge"), kern, mode="valid") for s in spec])
    height = rng.random((K, C2))
    img1 = spec[gt] + 0.05 * rng.standard_normal((H, W, C1))
    img2 = height[gt] + 0.05 * rng.standard_normal((H, W, C2))
    return minmax_normalise(img1), minmax_normalise(img2), gt

"""Host-side mirror of the reference's window utilities (utils.py:357-415, 567-582) with the
same names and argument meaning, plus the vectorised forms the CUDA path consumes."""
from __future__ import annotations

import itertools

import numpy as np


def window_starts(extent: int, win: int, step: int) -> np.ndarray:
    """Start offsets along one axis, exactly the values utils.sliding_window visits
    (utils.py:374-397): ``range(0, extent-win+offset+1, step)`` with ``offset =
    (extent-win) % step`` and overshooting starts clamped to ``extent-win``."""
    if win > extent:
        return np.zeros((0,), dtype=np.int32)
    offset = (extent - win) % step
    s = np.arange(0, extent - win + offset + 1, step, dtype=np.int64)
    s[s + win > extent] = extent - win
    return s.astype(np.int32)


def sliding_window(image1, image2, step=10, window_size=(20, 20), with_data=True):
    """Generator with the reference's signature and order (utils.py:357-401)."""
    w, h = window_size
    for x in window_starts(image1.shape[0], w, step):
        for y in window_starts(image1.shape[1], h, step):
            x, y = int(x), int(y)
            if with_data:
                yield image1[x:x + w, y:y + h], image2[x:x + w, y:y + h], x, y, w, h
            else:
                yield x, y, w, h


def count_sliding_window(top, top2, step=10, window_size=(20, 20)):
    """utils.py:404-415, in closed form."""
    w, h = window_size
    return int(len(window_starts(top.shape[0], w, step)) * len(window_starts(top.shape[1], h, step)))


def grouper(n, iterable):
    """utils.py:567-582."""
    it = iter(iterable)
    while True:
        chunk = tuple(itertools.islice(it, n))
        if not chunk:
            return
        yield chunk


def row_band_ranges(n_rows: int, n_cols: int, world_size: int):
    """Split the window rows into ``world_size`` contiguous bands; returns (first_window,
    n_windows) per rank in the reference's row-major window order.  Rank r needs raster rows
    [x_first, x_last + P) only: its band plus a halo of P//2 rows on each side."""
    base, rem = divmod(n_rows, world_size)
    out, r0 = [], 0
    for r in range(world_size):
        rows = base + (1 if r < rem else 0)
        out.append((r0 * n_cols, rows * n_cols))
        r0 += rows
    return out


def band_geometry(H: int, W: int, P: int, stride: int, rank: int, world: int) -> dict:
    """What rank ``rank`` of ``world`` needs for its row band of the scene:
    ``first`` / ``count`` = its windows in the reference's order, ``x0:x1`` = raster rows to
    upload (band + halo), ``o0:o1`` = map rows it owns (window centres), ``xs`` = window-row
    starts relative to ``x0``.  ``count == 0`` when there are more ranks than window rows."""
    xs, ys = window_starts(H, P, stride), window_starts(W, P, stride)
    first, count = row_band_ranges(len(xs), len(ys), world)[rank]
    if count == 0:
        return dict(first=first, count=0, x0=0, x1=0, o0=0, o1=0, xs=xs[:0], ys=ys)
    r0, r1 = first // len(ys), (first + count - 1) // len(ys)
    x0, x1 = int(xs[r0]), int(xs[r1]) + P
    return dict(first=first, count=count, x0=x0, x1=x1, o0=x0 + P // 2, o1=x1 - P + P // 2 + 1,
                xs=xs[r0:r1 + 1] - x0, ys=ys)


def subband_bounds(nrows: int, P: int, pipeline: int, block: int = 0, depth: int = 0, min_rows: int = 21,
                   lead_small: bool = False) -> list:
    """Cut the ``nrows`` contiguous (stride 1) window rows of a band into at most ``pipeline`` sub-bands for
    the upload / compute / download pipeline of predict_scene_host; returns the bounds [0, ..., nrows].
    ``block`` / ``depth`` = the scene-block edge B and sharing depth D of the library's shared stem
    (csrc/vc_common.cuh blk_step / blk_count): n block rows are exact for B + (B - 2D)(n - 1) raster rows =
    that many - (P - 1) window rows, so sub-bands are sized to whole block rows (a cut elsewhere makes both
    neighbours recompute a block row) and handed out in ascending size; ``lead_small`` makes the first one
    the smallest allowed (a shorter first upload, the one nothing hides; measured: no gain).  ``block == 0``
    (no shared stem): equal sub-bands of >= ``min_rows`` rows."""
    if nrows <= 0:
        return [0, 0]
    step = block - 2 * depth
    cap1 = block - (P - 1)                       # window rows one block row serves
    if block <= 0 or depth <= 0 or step <= 0 or cap1 <= 0:
        nsub = max(1, min(int(pipeline), nrows // min_rows))
        return [(nrows * k) // nsub for k in range(nsub + 1)]
    n_min = 1 + max(0, -(-(min_rows - cap1) // step))          # block rows of the smallest sub-band allowed
    total_1 = 1 + max(0, -(-(nrows - cap1) // step))           # block rows of the band cut nowhere
    nsub = max(1, min(int(pipeline), total_1 // n_min))
    total = max(nsub * n_min, -(-(nrows - nsub * (cap1 - step)) // step))
    if lead_small and nsub > 1 and total - n_min >= (nsub - 1) * n_min:
        rest, m = total - n_min, nsub - 1        # the first sub-band is the smallest allowed, the others share the rest
        per = [n_min] + [rest // m + (1 if k >= m - rest % m else 0) for k in range(m)]
    else:
        per = [total // nsub + (1 if k >= nsub - total % nsub else 0) for k in range(nsub)]      # ascending
    bounds, left = [0], nrows
    for k, n in enumerate(per):
        later = sum(cap1 + step * (m - 1) for m in per[k + 1:])
        take = min(cap1 + step * (n - 1), left)
        take = max(take if k < nsub - 1 else left, left - later)
        bounds.append(bounds[-1] + take)
        left -= take
    return bounds


def plan_subbands(nrows: int, P: int, nmax: int, sync: bool, depth_of, block_of, lead_small: bool = False):
    """Sub-bands (window-row spans ``[a, b)``, possibly overlapping) of a dense band for predict_scene_host, or None
    when no candidate has a shared stem.  The library picks the scene-block edge (31 / 63 / 95) per call from the
    raster it is handed (``block_of(raster_rows, depth)``; ``depth_of(window_rows)`` = the sharing depth it will use), so
    the number of sub-bands decides how much stem work the cuts add: a Houston scene in 4 sub-bands of 85 window rows is
    exactly 4 block rows of 95 (what the uncut scene needs), in 6 it is 12 block rows of 63 (twice the stem).  For every
    count up to ``nmax`` the band is cut at whole block rows (subband_bounds); a sub-band a few rows short of one block
    row starts earlier instead of falling back to a smaller block edge (the overlapping window rows are simply computed
    and downloaded twice, bit-identically); cost = stem rows + the first upload when nothing hides it (``sync``) + a
    fixed cost per sub-band, in units of one uploaded raster row (Houston: 9.2 ms per 384 stem rows, 7 ms per 349
    uploaded rows, ~0.15 ms of launches per sub-band).  Returns (cost, spans)."""
    best = None
    for n in range(1, max(1, int(nmax)) + 1):
        rows_n = -(-nrows // n)
        depth = depth_of(rows_n)
        if depth <= 0:
            continue
        block = block_of(rows_n + P - 1, depth)
        cb = subband_bounds(nrows, P, n, block, depth, lead_small=lead_small)
        cand, stem_rows = [], 0
        for a, b in zip(cb[:-1], cb[1:]):
            if b > a:
                short = block - (b - a + P - 1)
                if 0 < short <= 8 and a - short >= 0:
                    a -= short
                h = b - a + P - 1
                bk = block_of(h, depth)
                step = bk - 2 * depth
                stem_rows += ((h - bk + step - 1) // step + 1) * (bk + 1)
                cand.append((a, b))
        if not cand:
            continue
        first = cand[0][1] - cand[0][0] + P - 1
        cost = 1.2 * stem_rows + (first if sync else 0) + 8 * len(cand)
        if best is None or cost < best[0]:
            best = (cost, cand)
    return best


def camel_to_snake(name: str) -> str:
    """utils.py:883-885: ``ViTCNN`` -> ``vi_tcnn``-style folder name of the checkpoint path (the two regular
    expressions are the identifier-splitting idiom the reference uses; the result must match for
    ``main.py --restore`` paths to line up)."""
    import re
    head = re.sub("(.)([A-Z][a-z]+)", r"\1_\2", name)
    return re.sub("([a-z0-9])([A-Z])", r"\1_\2", head).lower()


def seed_torch(seed: int = 1029) -> None:
    """utils.py:887-895: seed Python, numpy and torch RNGs (the determinism contract of the
    sample shuffle, datasets.py:506, and of weight initialisation)."""
    import os
    import random
    import torch
    random.seed(seed)
    os.environ["PYTHONHASHSEED"] = str(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed(seed)
        torch.cuda.manual_seed_all(seed)
    torch.backends.cudnn.benchmark = False
    torch.backends.cudnn.deterministic = True


def confusion_matrix(prediction, target, n_classes: int, ignored_labels=()):
    """K x K confusion matrix (rows = target) on the device: what metrics() feeds to sklearn
    (utils.py:595-611).  prediction / target: CUDA integer tensors of equal numel."""
    import torch
    from . import _lib
    if not (prediction.is_cuda and target.is_cuda):
        raise RuntimeError("confusion_matrix needs CUDA tensors (no CPU path)")
    prediction, target = prediction.contiguous(), target.contiguous()
    if prediction.numel() != target.numel():
        raise ValueError("prediction and target differ in size")
    mask = 0
    for l in ignored_labels:
        if 0 <= int(l) < 64:
            mask |= 1 << int(l)
    cm = torch.zeros(n_classes, n_classes, dtype=torch.int64, device=target.device)
    with torch.cuda.device(target.device):
        _lib.check(_lib.lib().vc_confusion_matrix(prediction.data_ptr(), prediction.element_size(), target.data_ptr(),
                                                  target.element_size(), target.numel(), n_classes, mask, cm.data_ptr(),
                                                  torch.cuda.current_stream().cuda_stream), "vc_confusion_matrix")
    return cm


def metrics(prediction, target, ignored_labels=[], n_classes=None):
    """utils.py:585-663 with the same keys and formulas; the confusion matrix is counted on the
    device (CUDA tensors in), the K x K arithmetic stays on the host in float64 like the
    reference's.  Returns the reference's dict: "Confusion matrix", "Accuracy", "F1 scores",
    "Precisions", "AA", "Kappa"."""
    import torch
    if n_classes is None:
        keep = torch.ones_like(target, dtype=torch.bool)
        for l in ignored_labels:
            keep &= target != l
        n_classes = int(target[keep].max().item()) + 1
    cm = confusion_matrix(prediction, target, n_classes, ignored_labels).cpu().numpy()
    results = {"Confusion matrix": cm}
    total = np.sum(cm)
    results["Accuracy"] = sum(cm[x][x] for x in range(len(cm))) * (100 / float(total))
    with np.errstate(divide="ignore", invalid="ignore"):
        rs, cs, dg = cm.sum(1), cm.sum(0), np.diag(cm)
        results["F1 scores"] = 2.0 * dg / (rs + cs)
        results["Precisions"] = 1.0 * dg / rs
        rec = dg / rs
        results["AA"] = np.mean(rec[~np.isnan(rec)])
        pa = np.trace(cm) / float(total)
        pe = np.sum(cs * rs) / float(total * total)
        results["Kappa"] = (pa - pe) / (1 - pe)
    return results


def minmax_normalise_(img, per_band: bool = True):
    """Min-max normalisation to [0, 1] of a CUDA raster f32 [H, W, C], in place (datasets.py:124-133:
    per band for the HSI cube; ``per_band=False`` = one min / max for the whole array, as the reference
    does for the LiDAR raster).  Bit-exact with numpy's float32 arithmetic."""
    import torch
    from . import _lib
    if not img.is_cuda or img.dtype != torch.float32 or not img.is_contiguous():
        raise RuntimeError("minmax_normalise_ needs a contiguous CUDA float32 raster (no CPU path)")
    C = img.shape[-1]
    scratch = torch.empty(2 * C, dtype=torch.float32, device=img.device)
    with torch.cuda.device(img.device):
        _lib.check(_lib.lib().vc_minmax_normalise(img.data_ptr(), img.numel() // C, C, 1 if per_band else 0,
                                                  scratch.data_ptr(), torch.cuda.current_stream().cuda_stream),
                   "vc_minmax_normalise")
    return img


def _fixed_number_split(sample_num: int, labels: np.ndarray, seed: int):
    """Flat train / test index lists with ``sample_num`` training pixels per class, drawn exactly like
    the reference's samplingFixedNum (utils.py:754-773): ``np.random.seed(seed)``, one shuffle of every
    class's pixel list in class order, then one shuffle of each concatenated list."""
    np.random.seed(seed)
    flat = labels.ravel()
    train, test = [], []
    for c in range(1, int(flat.max()) + 1 if flat.size else 1):
        idx = np.flatnonzero(flat == c).tolist()
        np.random.shuffle(idx)
        train += idx[:sample_num]
        test += idx[sample_num:]
    np.random.shuffle(train)
    np.random.shuffle(test)
    return train, test


def sample_gt(gt, train_size, mode="random", seed=0):
    """Split a 2-D label map into (train_gt, test_gt) -- the reference's sample_gt (utils.py:775-846):

    * ``random``: stratified ``sklearn.model_selection.train_test_split`` over the labelled pixels (the same
      third-party call, so the same draw under the same numpy RNG state); ``train_size`` > 1 is a count;
    * ``fixed``: per class an unstratified ``train_test_split`` of that class's pixels;
    * ``disjoint``: per class the rows above the first row where 90 % of ``train_size`` of the class's pixels
      lie above go to the test side;
    * ``random_fixednumber``: ``train_size`` pixels per class (`_fixed_number_split`).

    Upstream bugs not reproduced: ``fixed`` indexes with a list of lists (an error on numpy >= 1.23; the intent
    -- a (rows, cols) pair -- is used here) and ``random_fixednumber`` uses the removed ``np.int`` (utils.py:836)."""
    gt = np.asarray(gt)
    if gt.ndim != 2:
        raise ValueError("gt must be a 2-D label map")
    rows, cols = np.nonzero(gt)
    train_gt, test_gt = np.zeros_like(gt), np.zeros_like(gt)
    if train_size > 1:
        train_size = int(train_size)
    if mode in ("random", "fixed"):
        try:
            from sklearn.model_selection import train_test_split
        except ImportError as e:     # the reference has the same dependency
            raise RuntimeError("sample_gt modes 'random' / 'fixed' need scikit-learn") from e
    if mode == "random":
        X = list(zip(rows, cols))
        tr, te = train_test_split(X, train_size=train_size, stratify=gt[rows, cols].ravel())
        tr, te = np.asarray(tr).reshape(-1, 2), np.asarray(te).reshape(-1, 2)
        train_gt[tr[:, 0], tr[:, 1]] = gt[tr[:, 0], tr[:, 1]]
        test_gt[te[:, 0], te[:, 1]] = gt[te[:, 0], te[:, 1]]
    elif mode == "fixed":
        for c in np.unique(gt):
            if c == 0:
                continue
            X = list(zip(*np.nonzero(gt == c)))
            tr, te = train_test_split(X, train_size=train_size)
            tr, te = np.asarray(tr).reshape(-1, 2), np.asarray(te).reshape(-1, 2)
            train_gt[tr[:, 0], tr[:, 1]] = c
            test_gt[te[:, 0], te[:, 1]] = c
    elif mode == "disjoint":
        train_gt, test_gt = np.copy(gt), np.copy(gt)
        for c in np.unique(gt):
            mask = gt == c
            total = int(mask.sum())
            above = np.concatenate([[0], np.cumsum(mask.sum(axis=1))])       # pixels of the class in rows [0, x)
            x = gt.shape[0] - 1
            for r in range(gt.shape[0]):
                if total and above[r] / total > 0.9 * train_size:
                    x = r
                    break
            mask[:x, :] = False
            train_gt[mask] = 0
        test_gt[train_gt > 0] = 0
    elif mode == "random_fixednumber":
        flat = gt.reshape(-1).astype(np.int64)
        tr, te = _fixed_number_split(int(train_size), flat, seed)
        train_flat, test_flat = np.zeros(flat.shape), np.zeros(flat.shape)      # float64, like the reference
        train_flat[tr] = flat[tr]
        test_flat[te] = flat[te]
        train_gt, test_gt = train_flat.reshape(gt.shape), test_flat.reshape(gt.shape)
    else:
        raise ValueError("{} sampling is not implemented yet.".format(mode))
    return train_gt, test_gt

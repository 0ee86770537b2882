"""Golden train / test splits produced by the reference's own utils.sample_gt / samplingFixedNum
(utils.py:754-846), imported with the stub recipe of SURVEY App. B.  Run in the build container
(/root/reference mounted): python tests/golden/make_split_golden.py"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def label_map(seed=3, shape=(40, 37), k=6):
    rng = np.random.default_rng(seed)
    gt = rng.integers(0, k, size=shape).astype(np.uint8)
    gt[rng.random(shape) < 0.3] = 0
    return gt


def main():
    for n in ["seaborn", "spectral", "visdom", "matplotlib", "matplotlib.pyplot"]:
        sys.modules.setdefault(n, types.ModuleType(n))
    sys.path.insert(0, "/root/reference")
    import utils as R
    gt = label_map()
    out = {"gt": gt}
    for tag, mode, ts in [("random_frac", "random", 0.2), ("random_count", "random", 50), ("disjoint_50", "disjoint", 0.5),
                          ("disjoint_30", "disjoint", 0.3)]:
        np.random.seed(7)
        tr, te = R.sample_gt(gt, ts, mode=mode)
        out[tag + "_train"], out[tag + "_test"] = tr, te
    tr, te = R.samplingFixedNum(5, gt.reshape(-1).astype(np.int64), 11)
    out["fixednum_train_idx"], out["fixednum_test_idx"] = np.asarray(tr), np.asarray(te)
    np.savez_compressed(os.path.join(HERE, "split_golden.npz"), **out)
    print("wrote split_golden.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()

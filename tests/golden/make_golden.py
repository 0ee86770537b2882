"""Generate the committed golden fixtures by running the REFERENCE's own code.

Run in the dev container (needs /root/reference):  python tests/golden/make_golden.py

Outputs (small, committed):
  data_golden.npz   - utils.sliding_window / count_sliding_window / grouper,
                      datasets.MultiModalX (indices after the seeded shuffle, samples,
                      flip/rotate augmentation), model_utils.test() and val() driven by
                      a deterministic toy network, utils.metrics().
  model_golden.npz  - logits of the in-repo fp32 oracle model (oracle/model_ref.py) for
                      seeded weights and inputs.  The reference ships no model source, so
                      this file only freezes the reconstruction (PARITY UNPINNED).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_import  # noqa: E402
from oracle.model_ref import ViTCNNRef, randomize_bn_stats  # noqa: E402


class ToyNet(torch.nn.Module):
    """Deterministic stand-in network: logits are fixed linear functionals of the
    patch, so test()/val() goldens pin gather order + scatter, not a model."""

    def __init__(self, K):
        super().__init__()
        self.K = K

    def forward(self, hsi, lidar):
        P = hsi.shape[-1]
        c = P // 2
        ks = torch.arange(1, self.K + 1, dtype=torch.float32)
        centre = hsi[:, :, c, c].sum(1)
        corner = hsi[:, 0, 0, 0] - hsi[:, -1, -1, -1]
        lid = lidar.reshape(lidar.shape[0], -1).mean(1)
        return centre[:, None] * ks[None] + corner[:, None] * ks[None].flip(1) + lid[:, None]


def toy_net_numpy(hsi, lidar, K):
    P = hsi.shape[-1]
    c = P // 2
    ks = np.arange(1, K + 1, dtype=np.float32)
    centre = hsi[:, :, c, c].sum(1, dtype=np.float32)
    corner = hsi[:, 0, 0, 0] - hsi[:, -1, -1, -1]
    lid = lidar.reshape(lidar.shape[0], -1).mean(1, dtype=np.float32)
    return centre[:, None] * ks[None] + corner[:, None] * ks[None][:, ::-1] + lid[:, None]


def main():
    ref_utils, ref_datasets, ref_mu = ref_import.import_reference()
    out = {}

    # ---- 1. sliding window enumeration -------------------------------------------
    sw_cases = []
    for (H, W) in [(20, 23), (17, 31), (12, 12), (11, 40), (15, 16)]:
        for P in (1, 5, 7, 8, 9, 11):
            for step in (1, 2, 3):
                if P > H or P > W:
                    continue
                a = np.zeros((H, W, 1), np.float32)
                corners = np.array([(x, y) for x, y, w, h in
                                    ref_utils.sliding_window(a, a, step=step, window_size=(P, P),
                                                             with_data=False)], dtype=np.int64).reshape(-1, 2)
                n = ref_utils.count_sliding_window(a, a, step=step, window_size=(P, P))
                assert n == len(corners)
                key = f"sw_{H}_{W}_{P}_{step}"
                out[key] = corners
                sw_cases.append((H, W, P, step))
    out["sw_cases"] = np.array(sw_cases, dtype=np.int64)
    out["grouper_7_of_23"] = np.array([len(c) for c in ref_utils.grouper(7, range(23))], dtype=np.int64)

    # ---- 2. MultiModalX ---------------------------------------------------------------
    rng = np.random.default_rng(123)
    ds_cases = []
    for ci, (H, W, C1, C2, P) in enumerate([(26, 31, 6, 1, 5), (26, 31, 6, 2, 7), (30, 27, 5, 1, 9),
                                           (29, 33, 4, 1, 11), (16, 18, 3, 1, 8)]):
        img1 = rng.random((H, W, C1), dtype=np.float32)
        img2 = rng.random((H, W, C2), dtype=np.float32)
        gt = rng.integers(0, 4, size=(H, W)).astype(np.uint8)
        hp = dict(dataset="synthetic", patch_size=P, ignored_labels=[0], flip_augmentation=False,
                  radiation_augmentation=False, mixture_augmentation=False, center_pixel=True,
                  supervision="full", applyPCA=False)
        ref_utils.seed_torch(ci)
        ds = ref_datasets.MultiModalX(img1, img2, gt, **hp)
        n = min(len(ds), 12)
        samples = [ds[i] for i in range(n)]
        out[f"ds{ci}_img1"], out[f"ds{ci}_img2"], out[f"ds{ci}_gt"] = img1, img2, gt
        out[f"ds{ci}_indices"] = np.asarray(ds.indices, dtype=np.int64).reshape(-1, 2)
        out[f"ds{ci}_hsi"] = np.stack([s[0].numpy() for s in samples])
        out[f"ds{ci}_lidar"] = np.stack([s[1].numpy() for s in samples])
        out[f"ds{ci}_label"] = np.array([int(s[2]) for s in samples], dtype=np.int64)
        # flip / rotate augmentation: replay the same numpy seed so the oracle's
        # restatement of the decision draws can be compared to these outputs
        hp_aug = dict(hp, flip_augmentation=True)
        ref_utils.seed_torch(100 + ci)
        ds_aug = ref_datasets.MultiModalX(img1, img2, gt, **hp_aug)
        aug = [ds_aug[i] for i in range(n)]
        out[f"ds{ci}_aug_hsi"] = np.stack([s[0].numpy() for s in aug])
        out[f"ds{ci}_aug_lidar"] = np.stack([s[1].numpy() for s in aug])
        out[f"ds{ci}_aug_label"] = np.array([int(s[2]) for s in aug], dtype=np.int64)
        ds_cases.append((H, W, C1, C2, P, n))
    out["ds_cases"] = np.array(ds_cases, dtype=np.int64)

    # ---- 3. test(): gather order + logits scatter ----------------------------------
    K = 5
    net = ToyNet(K)
    t_cases = []
    for ti, (H, W, C1, C2, P, stride, bs) in enumerate([(19, 22, 4, 1, 5, 1, 7), (19, 22, 4, 2, 7, 2, 16),
                                                       (23, 17, 3, 1, 9, 3, 5), (14, 30, 6, 1, 11, 1, 64)]):
        img1 = rng.random((H, W, C1), dtype=np.float32)
        img2 = rng.random((H, W, C2), dtype=np.float32)
        hp = dict(patch_size=P, center_pixel=True, batch_size=bs, device=torch.device("cpu"),
                  n_classes=K, applyPCA=False, test_stride=stride)
        probs = ref_mu.test(0, net, img1, img2, hp)
        out[f"test{ti}_img1"], out[f"test{ti}_img2"], out[f"test{ti}_probs"] = img1, img2, probs
        t_cases.append((H, W, C1, C2, P, stride, bs))
        # numpy twin of the toy net must agree with the torch one to float32 bits
        chk = toy_net_numpy(img1[None, :P, :P].transpose(0, 3, 1, 2), img2[None, :P, :P].transpose(0, 3, 1, 2), K)
        ref = net(torch.from_numpy(np.ascontiguousarray(img1[None, :P, :P].transpose(0, 3, 1, 2))),
                  torch.from_numpy(np.ascontiguousarray(img2[None, :P, :P].transpose(0, 3, 1, 2)))).numpy()
        assert np.allclose(chk, ref, rtol=1e-5, atol=1e-5)
    out["test_cases"] = np.array(t_cases, dtype=np.int64)

    # ---- 4. val() ---------------------------------------------------------------------
    H, W, C1, C2, P = 22, 25, 4, 1, 5
    img1 = rng.random((H, W, C1), dtype=np.float32)
    img2 = rng.random((H, W, C2), dtype=np.float32)
    gt = rng.integers(0, K, size=(H, W)).astype(np.uint8)
    hp = dict(dataset="synthetic", patch_size=P, ignored_labels=[0], flip_augmentation=False,
              radiation_augmentation=False, mixture_augmentation=False, center_pixel=True,
              supervision="full", applyPCA=False)
    ref_utils.seed_torch(7)
    ds = ref_datasets.MultiModalX(img1, img2, gt, **hp)
    loader = torch.utils.data.DataLoader(ds, batch_size=16, shuffle=False)

    class ModNet(torch.nn.Module):
        def forward(self, hsi, lidar):  # predictions cycle through classes incl. ignored 0
            s = (hsi[:, 0, P // 2, P // 2] * 1000).long() % K
            return torch.nn.functional.one_hot(s, K).float()

    out["val_img1"], out["val_img2"], out["val_gt"] = img1, img2, gt
    out["val_indices"] = np.asarray(ds.indices, dtype=np.int64)
    out["val_acc"] = np.array(ref_mu.val(ModNet(), loader, device="cpu"))

    # ---- 5. metrics() -----------------------------------------------------------------
    pred = rng.integers(0, K, size=(40, 37))
    tgt = rng.integers(0, K, size=(40, 37))
    res = ref_utils.metrics(pred, tgt, ignored_labels=[0], n_classes=K)
    out["met_pred"], out["met_tgt"] = pred, tgt
    out["met_cm"] = np.asarray(res["Confusion matrix"], dtype=np.int64)
    out["met_acc"], out["met_f1"] = np.array(res["Accuracy"]), np.asarray(res["F1 scores"])
    out["met_prec"], out["met_aa"], out["met_kappa"] = np.asarray(res["Precisions"]), np.array(res["AA"]), np.array(res["Kappa"])

    np.savez_compressed(os.path.join(HERE, "data_golden.npz"), **out)
    print("data_golden.npz:", len(out), "arrays")

    # ---- 6. model oracle self-golden (PARITY UNPINNED) ---------------------------------
    mout = {}
    for name, (C1, C2, P, K_, B) in {"small": (16, 1, 5, 4, 3), "muufl7": (64, 2, 7, 12, 2),
                                      "houston11": (144, 1, 11, 16, 2)}.items():
        torch.manual_seed(0)
        m = ViTCNNRef(C1, C2, patch_size=P, num_classes=K_)
        randomize_bn_stats(m, seed=1)
        m.eval()
        g = torch.Generator().manual_seed(5)
        hsi = torch.rand(B, C1, P, P, generator=g)
        lid = torch.rand(B, C2, P, P, generator=g)
        with torch.no_grad():
            logits = m(hsi, lid)
            tok = m.tokens(hsi, lid)
        mout[f"{name}_cfg"] = np.array([C1, C2, P, K_, B], dtype=np.int64)
        mout[f"{name}_hsi"], mout[f"{name}_lidar"] = hsi.numpy(), lid.numpy()
        mout[f"{name}_logits"], mout[f"{name}_tokens"] = logits.numpy(), tok.numpy()
        if name != "houston11":   # keep the fixture small; houston weights are re-seeded
            for k, v in m.state_dict().items():
                mout[f"{name}_sd_{k}"] = v.numpy()
    np.savez_compressed(os.path.join(HERE, "model_golden.npz"), **mout)
    print("model_golden.npz:", len(mout), "arrays")


if __name__ == "__main__":
    main()

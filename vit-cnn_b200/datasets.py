"""Device-resident mirror of the reference's training dataset (datasets.py:461-593,
``MultiModalX``): same constructor arguments, sample set, shuffle draw and per-item values, with
the rasters held in HBM and patches cut by the gather kernel (bit-exact fp32 copies).

The per-item protocol (``__len__`` / ``__getitem__``) is kept for ``torch.utils.data.DataLoader``
compatibility; ``batch()`` / ``loader()`` are the fast path (one launch per batch).
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops


class MultiModalX(torch.utils.data.Dataset):
    def __init__(self, data, data2, gt, **hyperparams):
        super().__init__()
        self.name = hyperparams["dataset"]
        self.patch_size = hyperparams["patch_size"]
        self.ignored_labels = set(hyperparams["ignored_labels"])
        self.flip_augmentation = hyperparams["flip_augmentation"]
        self.radiation_augmentation = hyperparams["radiation_augmentation"]
        self.mixture_augmentation = hyperparams["mixture_augmentation"]
        self.center_pixel = hyperparams["center_pixel"]
        if self.mixture_augmentation:
            raise NotImplementedError("mixture_augmentation is broken upstream (its label lookup compares a list with "
                                      "a scalar and raises, SURVEY.md App. C rule 2) and is not implemented")
        if hyperparams.get("applyPCA", False) == True:  # noqa: E712
            raise ValueError("ViT-CNN runs on the full band set: applyPCA must be False")
        if not self.center_pixel or self.patch_size < 2:
            raise ValueError("ViT-CNN classifies the centre pixel of a P x P patch (center_pixel=True, P > 1)")
        supervision = hyperparams["supervision"]
        gt = np.asarray(gt)
        if supervision == "full":                       # datasets.py:488-492
            mask = np.ones_like(gt)
            for l in self.ignored_labels:
                mask[gt == l] = 0
        elif supervision == "semi":                     # :494-495
            mask = np.ones_like(gt)
        else:
            raise ValueError('supervision mode "{}" is unknown.'.format(supervision))
        x_pos, y_pos = np.nonzero(mask)
        p = self.patch_size // 2
        H, W = data.shape[0], data.shape[1]
        keep = (x_pos > p) & (x_pos < H - p) & (y_pos > p) & (y_pos < W - p)      # :497-504, strict
        self.indices = np.stack([x_pos[keep], y_pos[keep]], axis=1) if keep.any() else np.array([])
        self.labels = [gt[x, y] for x, y in self.indices]                        # :505 (pre-shuffle order)
        np.random.shuffle(self.indices)                                         # :506
        device = torch.device(hyperparams.get("device", "cuda"))
        if device.type != "cuda":
            raise RuntimeError("the device-resident dataset needs a CUDA device (no CPU path)")
        self.device = device
        self.data = torch.as_tensor(np.ascontiguousarray(data, dtype=np.float32)).to(device)
        self.data2 = torch.as_tensor(np.ascontiguousarray(data2, dtype=np.float32)).to(device)
        self.label = torch.as_tensor(np.ascontiguousarray(gt).astype(np.int64)).to(device)
        self._xy = torch.as_tensor(np.asarray(self.indices, dtype=np.int32).reshape(-1, 2)).to(device)

    def __len__(self):
        return len(self.indices)

    def draw_augmentation(self, n: int):
        """The numpy RNG draws of n consecutive ``__getitem__`` calls (datasets.py:559-566, 510-532), in
        the reference's order: per item the flip / rotate decision, then the radiation-noise decision with
        its alpha and its [P,P,C1] normal noise.  Returns (ops uint8 [n], list of (item, alpha, noise))."""
        P, C1 = self.patch_size, self.data.shape[2]
        codes, rad = np.zeros(n, dtype=np.uint8), []
        for i in range(n):
            if self.flip_augmentation and P > 1:
                if np.random.random() > 0.5:                       # flip(): horizontal, then vertical draw
                    horizontal = np.random.random() > 0.5
                    vertical = np.random.random() > 0.5
                    codes[i] = int(horizontal) + 2 * int(vertical)
                elif np.random.random() > 0.5:                     # rotate()
                    codes[i] = 3 + int(np.random.choice([1, 2, 3]))
            if self.radiation_augmentation and np.random.random() < 0.1:
                alpha = np.random.uniform(0.9, 1.1)
                rad.append((i, alpha, np.random.normal(loc=0.0, scale=1.0, size=(P, P, C1))))
        return codes, rad

    def batch(self, idx):
        """(data [B,C1,P,P] f32, data2 [B,C2,P,P] f32, target int64 [B]) for sample numbers idx; with the
        augmentation flags on, the RNG decisions are drawn on the host exactly as the reference's
        ``__getitem__`` would for these items in this order, flips / rot90 are applied by the gather
        kernel (index remap), radiation noise as alpha * x (float32) + noise / 25 (float64) like numpy."""
        idx = torch.as_tensor(idx, dtype=torch.int64, device=self.device).reshape(-1)
        xy = self._xy[idx].contiguous()
        codes, rad = None, []
        if self.flip_augmentation or self.radiation_augmentation:
            c, rad = self.draw_augmentation(idx.numel())
            codes = torch.from_numpy(c).to(self.device) if c.any() else None
        data, data2, target = ops.gather_patches(self.data, self.data2, xy, self.patch_size, center_mode=True,
                                                 gt=self.label, ops=codes, validate=False)
        for i, alpha, noise in rad:                                   # datasets.py:528-532
            nz = torch.from_numpy(noise).to(self.device).permute(2, 0, 1)
            # numpy: python-float alpha times a float32 array stays float32; the float64 noise term promotes the sum
            data[i] = ((data[i] * float(alpha)).double() + (1 / 25) * nz).float()
        return data, data2, target

    def centres(self, idx):
        idx = torch.as_tensor(idx, dtype=torch.int64, device=self.device).reshape(-1)
        return self._xy[idx].contiguous()

    def __getitem__(self, i):
        d, d2, t = self.batch([int(i)])
        return d[0], d2[0], t[0]

    def loader(self, batch_size, shuffle=False, generator=None):
        return _Loader(self, batch_size, shuffle, generator)


def batch_indices(n: int, batch_size: int, shuffle: bool, generator=None):
    """Index batches in exactly the order ``torch.utils.data.DataLoader(dataset_of_n, batch_size, shuffle)``
    visits them under the current torch RNG state (main.py:434-447: no drop_last, num_workers=0): the order is
    produced by torch's own DataLoader over the index range, so the draws from the global generator (the
    iterator's base seed, then the RandomSampler's seed) are the same calls in the same order."""
    index_loader = torch.utils.data.DataLoader(range(n), batch_size=int(batch_size), shuffle=bool(shuffle),
                                               generator=generator, num_workers=0,
                                               collate_fn=lambda items: torch.as_tensor(items, dtype=torch.int64))
    return iter(index_loader)


class _Loader:
    """DataLoader stand-in (what train() / val() touch: iteration, len(), .dataset) that cuts one batch per
    launch of the gather kernel; same batches in the same order as the stock DataLoader over this dataset."""

    def __init__(self, dataset, batch_size, shuffle, generator):
        self.dataset, self.batch_size, self.shuffle, self.generator = dataset, int(batch_size), shuffle, generator

    def __len__(self):
        return (len(self.dataset) + self.batch_size - 1) // self.batch_size      # no drop_last (main.py:434-447)

    def __iter__(self):
        for idx in batch_indices(len(self.dataset), self.batch_size, self.shuffle, self.generator):
            yield self.dataset.batch(idx)

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
        if self.flip_augmentation or self.radiation_augmentation or self.mixture_augmentation:
            raise NotImplementedError("augmentations are not implemented on the device path yet "
                                      "(mixture_augmentation is broken upstream: SURVEY.md App. C rule 2)")
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

    def batch(self, idx):
        """(data [B,C1,P,P] f32, data2 [B,C2,P,P] f32, target int64 [B]) for sample numbers idx."""
        idx = torch.as_tensor(idx, dtype=torch.int64, device=self.device).reshape(-1)
        xy = self._xy[idx].contiguous()
        return ops.gather_patches(self.data, self.data2, xy, self.patch_size, center_mode=True, gt=self.label)

    def centres(self, idx):
        idx = torch.as_tensor(idx, dtype=torch.int64, device=self.device).reshape(-1)
        return self._xy[idx].contiguous()

    def __getitem__(self, i):
        d, d2, t = self.batch([int(i)])
        return d[0], d2[0], t[0]

    def loader(self, batch_size, shuffle=False, generator=None):
        return _Loader(self, batch_size, shuffle, generator)


class _Loader:
    """Minimal DataLoader stand-in (what train() / val() touch: iteration, len(), .dataset)."""

    def __init__(self, dataset, batch_size, shuffle, generator):
        self.dataset, self.batch_size, self.shuffle, self.generator = dataset, int(batch_size), shuffle, generator

    def __len__(self):
        return (len(self.dataset) + self.batch_size - 1) // self.batch_size      # no drop_last (main.py:434-447)

    def __iter__(self):
        n = len(self.dataset)
        order = torch.randperm(n, generator=self.generator) if self.shuffle else torch.arange(n)
        for s in range(0, n, self.batch_size):
            yield self.dataset.batch(order[s:s + self.batch_size])

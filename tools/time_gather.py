"""Roofline of the exact fp32 patch gather (vc_gather_patches_f32 = MultiModalX.__getitem__ + collate, datasets.py:550-593;
test()'s batch assembly, model_utils.py:1103-1112): algorithmic bytes = 2 x (C1 + C2) x P^2 x 4 per window (read + write,
SURVEY.md section 8(d): 140 360 B at the Houston shape) against the measured HBM copy bandwidth."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import vitcnn_b200  # noqa: F401
from vitcnn_b200 import ops

H, W, C1, C2, P = 349, 1905, 144, 1, 11
n = int(os.environ.get("N", 65536))
dev = "cuda:0"
g = torch.Generator(device=dev).manual_seed(0)
img1 = torch.rand(H, W, C1, device=dev, generator=g)
img2 = torch.rand(H, W, C2, device=dev, generator=g)
peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.isfile(
    os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else 6650.0
for mode in ("dense", "random"):
    if mode == "dense":          # consecutive windows of the sliding-window order (test())
        k = torch.arange(n, device=dev)
        xy = torch.stack([k // (W - P + 1) + 100, k % (W - P + 1)], 1).to(torch.int32)
        center = False
    else:                        # shuffled labelled pixels (training batches)
        xy = torch.stack([torch.randint(P // 2 + 1, H - P // 2 - 1, (n,), device=dev, generator=g),
                          torch.randint(P // 2 + 1, W - P // 2 - 1, (n,), device=dev, generator=g)], 1).to(torch.int32)
        center = True
    for _ in range(3):
        out = ops.gather_patches(img1, img2, xy, P, center_mode=center, validate=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        out = ops.gather_patches(img1, img2, xy, P, center_mode=center, validate=False)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    alg = 2.0 * (C1 + C2) * P * P * 4 * n
    print(json.dumps({"kernel": "gather_f32_kernel (HSI + LiDAR launches) incl. output allocation", "windows": n, "order": mode,
                      "ms": round(ms, 3), "algorithmic_bytes_per_window": 2 * (C1 + C2) * P * P * 4,
                      "achieved_GBps": round(alg / ms / 1e6, 1), "peak_GBps": peak, "frac": round(alg / ms / 1e6 / peak, 3)}))

#!/usr/bin/env python
"""BASELINE.json configs[3] / [4] on one GPU: full-scene inference (pixels/s) and one training step at global batch
4096 (samples/s) on synthetic MUUFL- and Augsburg-shaped rasters and for the patch-size sweep 7 / 9 / 11 / 15 at the
Houston shape.  Device-resident inputs, CUDA events, 3 warm-up + 5 timed repetitions; one JSON line per case."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vitcnn_b200
from vitcnn_b200 import _lib
from vitcnn_b200.train import Trainer

CASES = [("muufl_p11", 325, 220, 64, 2, 12, 11), ("augsburg_p11", 332, 485, 180, 1, 8, 11),
         ("houston_p7", 349, 1905, 144, 1, 16, 7), ("houston_p9", 349, 1905, 144, 1, 16, 9),
         ("houston_p11", 349, 1905, 144, 1, 16, 11), ("houston_p15", 349, 1905, 144, 1, 16, 15)]
dev = torch.device("cuda:0")


def timed(fn, warm=3, reps=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


only = set(sys.argv[1:])
for name, H, W, C1, C2, K, P in CASES:
    if only and name not in only:
        continue
    rng = np.random.default_rng(0)
    img1 = torch.from_numpy(rng.random((H, W, C1), dtype=np.float32)).to(dev)
    img2 = torch.from_numpy(rng.random((H, W, C2), dtype=np.float32)).to(dev)
    gt = torch.from_numpy(rng.integers(1, K, size=(H, W)).astype(np.int64)).to(dev)
    torch.manual_seed(0)
    net = vitcnn_b200.ViTCNN(C1, C2, patch_size=P, num_classes=K).to(dev).eval()
    lm = torch.zeros(H, W, K, device=dev)
    am = torch.zeros(H, W, dtype=torch.uint8, device=dev)
    ms_inf = timed(lambda: net.predict_scene(img1, img2, logits_map=lm, argmax_map=am))
    pk = net.pack_for_inference()
    import ctypes
    nwin = (H - P + 1) * (W - P + 1)
    depth = _lib.lib().vc_scene_shared_depth(ctypes.byref(pk["struct"]), H, W, min(32768, nwin), nwin, 1 << 40)
    line = {"case": name, "scene": [H, W, C1, C2], "classes": K, "patch": P, "infer_ms": round(ms_inf, 3),
            "pixels_per_s": H * W / (ms_inf / 1e3), "windows_per_s": nwin / (ms_inf / 1e3), "shared_stem_depth": int(depth)}
    try:
        tnet = vitcnn_b200.ViTCNN(C1, C2, patch_size=P, num_classes=K, dropout=0.01).to(dev)
        w = torch.ones(K)
        w[0] = 0
        tr = Trainer(tnet, lr=1e-3, weights=w, use_graph=True)
        p = P // 2
        xy = [torch.from_numpy(np.stack([rng.integers(p + 1, H - p - 1, 4096), rng.integers(p + 1, W - p - 1, 4096)], 1)
                               .astype(np.int32)).to(dev) for _ in range(4)]
        it = [0]

        def step():
            tr.step(img1, img2, gt, xy[it[0] % 4])
            it[0] += 1
        ms_tr = timed(step, warm=4, reps=10)
        line.update({"train_ms_per_step": round(ms_tr, 3), "train_samples_per_s": 4096 / (ms_tr / 1e3)})
    except Exception as e:      # a patch size the training kernels do not cover is reported, not hidden
        line["train_error"] = str(e)[:200]
    print(json.dumps(line), flush=True)
    del net, lm, am, img1, img2
    torch.cuda.empty_cache()

"""Chunk-size sweep of full-scene inference on the Houston shape: each chunk size must reproduce the
chunk = 32768 logits / argmax maps bit for bit (large chunks exercise the 64-bit workspace offsets), then
device-resident and host-buffer times per scene (CUDA events, 3 warm-up + 5 timed passes)."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vitcnn_b200  # noqa: E402

H, W, C1, C2, P, K = 349, 1905, 144, 1, 11, 16
dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
img1_h = torch.from_numpy(rng.random((H, W, C1), dtype=np.float32)).pin_memory()
img2_h = torch.from_numpy(rng.random((H, W, C2), dtype=np.float32)).pin_memory()
torch.manual_seed(0)
net = vitcnn_b200.ViTCNN(C1, C2, patch_size=P, num_classes=K).to(dev).eval()
img1, img2 = img1_h.to(dev), img2_h.to(dev)
lg_h = torch.empty(H, W, K).pin_memory()
am_h = torch.empty(H, W, dtype=torch.uint8).pin_memory()


def timed(fn, n=5, w=3):
    for _ in range(w):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


want_lg, want_am = net.predict_scene(img1, img2, chunk=32768)
want_lg, want_am = want_lg.clone(), want_am.clone()
for chunk in [int(c) for c in sys.argv[1:]] or [32768, 131072, 262144]:
    net._ws = None
    torch.cuda.empty_cache()
    lg, am = net.predict_scene(img1, img2, chunk=chunk)
    same = bool(torch.equal(lg, want_lg) and torch.equal(am, want_am))
    lg_h.zero_()
    am_h.zero_()
    vitcnn_b200.predict_scene_host(net, img1_h, img2_h, chunk=chunk, logits_out=lg_h, argmax_out=am_h)
    same_host = bool(torch.equal(lg_h, want_lg.cpu()) and torch.equal(am_h, want_am.cpu()))
    ms = timed(lambda: net.predict_scene(img1, img2, chunk=chunk))
    ms_host = timed(lambda: vitcnn_b200.predict_scene_host(net, img1_h, img2_h, chunk=chunk, logits_out=lg_h, argmax_out=am_h))
    print(json.dumps({"chunk": chunk, "bit_identical": same, "bit_identical_host": same_host, "device_ms": round(ms, 2),
                      "e2e_ms": round(ms_host, 2), "workspace_GB": round(net._ws.numel() / 1e9, 2),
                      "peak_alloc_GB": round(torch.cuda.max_memory_allocated() / 1e9, 2)}), flush=True)

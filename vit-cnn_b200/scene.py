"""Host-buffer entry point of full-scene inference: upload the rank's row band (plus halo
rows) from pinned host memory, run the scene kernels, download the band's logits / argmax.
This is the `e2e` path of bench.py and what test() uses; row bands need no collective
(SURVEY.md section 8(e))."""
from __future__ import annotations

import torch

from .utils import band_geometry


@torch.no_grad()
def predict_scene_host(net, img1: torch.Tensor, img2: torch.Tensor, stride: int = 1, rank: int = 0, world: int = 1,
                       chunk: int = 2048, logits_out: torch.Tensor = None, argmax_out: torch.Tensor = None,
                       device=None):
    """img1 f32 [H,W,C1], img2 f32 [H,W,C2] CPU tensors (pinned for async copies).  Writes the
    rows owned by ``rank`` into ``logits_out`` f32 [H,W,K] / ``argmax_out`` uint8 [H,W] (CPU,
    allocated zero-filled when None) and returns them.  Rows no window is centred on are not
    touched."""
    H, W, _ = img1.shape
    P, K = net.patch_size, net.num_classes
    dev = torch.device(device) if device is not None else net.cls_token.device
    if logits_out is None:
        logits_out = torch.zeros(H, W, K, dtype=torch.float32)
    if argmax_out is None:
        argmax_out = torch.zeros(H, W, dtype=torch.uint8)
    geo = band_geometry(H, W, P, stride, rank, world)
    if geo["count"] == 0:
        return logits_out, argmax_out
    x0, x1 = geo["x0"], geo["x1"]
    with torch.cuda.device(dev):
        b1 = img1[x0:x1].to(dev, non_blocking=True)
        b2 = img2[x0:x1].to(dev, non_blocking=True)
        lg, am = net.predict_scene(b1, b2, stride=stride, chunk=chunk, xs=geo["xs"])
        o0, o1 = geo["o0"], geo["o1"]                               # rows with window centres
        logits_out[o0:o1].copy_(lg[P // 2:P // 2 + (o1 - o0)], non_blocking=True)
        argmax_out[o0:o1].copy_(am[P // 2:P // 2 + (o1 - o0)], non_blocking=True)
        torch.cuda.current_stream().synchronize()
    return logits_out, argmax_out

"""Per-kernel timings (CUDA events) at Houston chunk shapes, with the conv kernel's bring-up
toggles, to locate bottlenecks.  usage: python tools/time_kernels.py [n_patches]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vitcnn_b200 import ops  # noqa: E402
from vitcnn_b200.model import pack_conv_weight, slices_for  # noqa: E402

dev = "cuda:0"
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
P = 11


def timeit(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def conv_case(cin, cout, ns, label, const=False):
    s_in = slices_for(cin)
    rows = ops.sps_rows(n, P)
    a = (torch.rand(s_in, rows, 8, device=dev) - 0.3).to(torch.bfloat16)
    w = pack_conv_weight(torch.randn(cout, cin, 3, 3) * 0.05, s_in, cout, ns).to(dev)
    if const:
        a = torch.full_like(a, 0.0078125)
        w = torch.full_like(w, 0.0078125)
        label += "-const"
    sc, bi = torch.ones(cout, device=dev), torch.zeros(cout, device=dev)
    out = torch.zeros(cout // 8, rows, 8, dtype=torch.bfloat16, device=dev)
    tiles = (n * (P + 1) ** 2 + 127) // 128
    mmas = -(-tiles // (148 // ns)) * 9 * (s_in // 2)
    for flags, name in [(0, "default"), (6, "no loads/stores"), (8, "no full waits"), (16, "no epilogue"),
                        (8 | 16, "no full, no epi"), (8 | 16 | 32, "issue only"), (1 << 12, "kpb 1"), (3 << 12, "kpb 3"), (4 << 12, "kpb 4")]:
        try:
            ms = timeit(lambda: ops.conv_sps(a, w, sc, bi, n, P, impl=0, debug_flags=flags, out=out))
        except RuntimeError as e:
            print(label, name, "failed", e)
            continue
        flops = 2.0 * n * 121 * cout * cin * 9
        print(f"{label:8s} {name:16s} {ms*1e3:8.1f} us  {flops/ms/1e9:8.1f} TFLOP/s(alg)  "
              f"{ms*1e-3*1.87e9/mmas:6.1f} cyc/MMA")


conv_case(144, 128, 2, "conv1")
conv_case(128, 64, 1, "conv2")
conv_case(64, 32, 1, "conv3")
conv_case(16, 16, 1, "lidar")

# pack (raster -> SPS) and the exact fp32 gather, Houston raster
H, W, C1 = 349, 1905, 144
img = torch.rand(H, W, C1, device=dev)
img2 = torch.rand(H, W, 1, device=dev)
xs = torch.arange(H - P + 1, dtype=torch.int32, device=dev)
ys = torch.arange(W - P + 1, dtype=torch.int32, device=dev)
off1, off2, oidx, xy = ops.scene_index(xs, ys, 100000, n, W, C1, 1, P)
ms = timeit(lambda: ops.pack_sps_raster(img, off1, P, slices_for(C1)))
print(f"pack_sps raster C=144: {ms*1e3:.1f} us, write {n*144*18*16/ms/1e6:.0f} GB/s, alg r+w {n*121*144*(4+2)/ms/1e6:.0f} GB/s")
for nn in (4096, 16384):
    ms = timeit(lambda: ops.gather_patches(img, img2, xy[:nn] if nn <= n else xy, P, center_mode=False))
    m = min(nn, n)
    print(f"gather_f32 n={m}: {ms*1e3:.1f} us, algorithmic r+w {2*m*145*121*4/ms/1e6:.0f} GB/s")

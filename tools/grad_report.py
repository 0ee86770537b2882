"""Per-parameter gradient comparison of the CUDA module vs the fp32 oracle (debug aid)."""
import sys
import torch
import torch.nn.functional as F
sys.path.insert(0, ".")
from tests.test_gpu_train import _pair, _grad_report, DEV

cfgs = [(16, 1, 5, 4, 6), (144, 1, 11, 16, 8), (144, 1, 11, 16, 64)]
for C1, C2, P, K, B in cfgs:
    ref, ours = _pair(C1, C2, P, K)
    g = torch.Generator().manual_seed(5)
    hsi, lid = torch.rand(B, C1, P, P, generator=g), torch.rand(B, C2, P, P, generator=g)
    y = torch.randint(1, K, (B,), generator=g)
    w = torch.ones(K); w[0] = 0
    F.cross_entropy(ref(hsi, lid), y, weight=w).backward()
    F.cross_entropy(ours(hsi.to(DEV), lid.to(DEV)), y.to(DEV), weight=w.to(DEV)).backward()
    print("cfg", (C1, C2, P, K, B))
    for k, e, c, s in _grad_report(ref, ours):
        print(f"  {k:34s} err {e:8.4f} cos {c:8.5f} scale {s:.3e}")

"""Bring-up probe for the tcgen05 conv kernel: one case, one debug flag, one process.
usage: python tools/probe_conv.py <case_index> <debug_flags>"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import emu  # noqa: E402
from tests.test_gpu_conv import CONV_CASES, _rand_conv  # noqa: E402
from vitcnn_b200 import ops  # noqa: E402
from vitcnn_b200.model import slices_for  # noqa: E402

ci, flags = int(sys.argv[1]), int(sys.argv[2])
cin, cout, n_out, ns, taps, P, n = CONV_CASES[ci]
s_in = slices_for(cin)
g = torch.Generator().manual_seed(1)
x = torch.rand(n, cin, P, P, generator=g) - 0.3
w, scale, bias = _rand_conv(cin, cout, s_in, n_out, ns, taps, seed=2)
a = emu.pack_sps(x, s_in)
want = emu.conv_sps(a, w, scale, bias, n, P, relu=True)
dev = "cuda:0"
ad, wd, sd, bd = a.to(torch.bfloat16).to(dev), w.to(dev), scale.to(dev), bias.to(dev)
ref = ops.conv_sps(ad, wd, sd, bd, n, P, impl=1).float().cpu()
torch.cuda.synchronize()
print(f"case {CONV_CASES[ci]} simt-vs-emu max err {(ref - want).abs().max().item():.4g}")
got = ops.conv_sps(ad, wd, sd, bd, n, P, impl=0, debug_flags=flags).float().cpu()
torch.cuda.synchronize()
err = (got - want).abs()
print(f"flags {flags}: tcgen05-vs-emu max err {err.max().item():.4g} (want max {want.abs().max().item():.4g}) "
      f"mismatch frac {(err > 2e-2).float().mean().item():.4f}")
if err.max().item() > 2e-2:
    bad = (err > 2e-2).nonzero()[:12]
    for s, r, k in bad.tolist():
        print("  slice", s, "row", r, "k", k, "got", got[s, r, k].item(), "want", want[s, r, k].item())
    H = emu.halo(P)
    rows = torch.arange(got.shape[1])
    per_row = err.amax((0, 2))
    print("  bad rows (first 40):", rows[per_row > 2e-2][:40].tolist())
    per_ch = err.permute(0, 2, 1).reshape(-1, got.shape[1]).amax(1)
    print("  bad channels:", (per_ch > 2e-2).nonzero().flatten()[:64].tolist())

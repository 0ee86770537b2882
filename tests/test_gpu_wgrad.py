"""-m gpu: tcgen05 weight-gradient kernel (rows = reduction axis, MN-major operands) vs
torch autograd on the same bf16-rounded operands."""
import pytest
import torch
import torch.nn.functional as F

from tests import emu

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

CASES = [  # cin, cout, taps, P, n, shift_on_a
    (144, 128, 9, 11, 5, False),
    (128, 64, 9, 11, 7, True),
    (64, 32, 9, 11, 9, True),
    (16, 16, 9, 5, 3, False),
    (180, 128, 9, 15, 3, False),
    (64, 128, 9, 7, 400, False),      # many tiles per CTA: pipeline wrap
    (32, 96, 1, 11, 6, False),
    (128, 32, 1, 7, 20, False),
]


@pytest.mark.parametrize("case", CASES)
def test_wgrad_vs_autograd(case):
    from vitcnn_b200 import ops
    from vitcnn_b200.model import slices_for
    cin, cout, taps, P, n, shift_on_a = case
    g = torch.Generator().manual_seed(cin + 3 * cout + P)
    x = emu.bf16(torch.rand(n, cin, P, P, generator=g) - 0.3)
    dy = emu.bf16(torch.randn(n, cout, P, P, generator=g))
    k = 3 if taps == 9 else 1
    w = torch.zeros(cout, cin, k, k, requires_grad=True)
    F.conv2d(x, w, padding=k // 2).backward(dy)
    want = w.grad
    xs = emu.pack_sps(x, slices_for(cin)).to(torch.bfloat16).to(DEV)
    dys = emu.pack_sps(dy, slices_for(cout)).to(torch.bfloat16).to(DEV)
    out = torch.full((cout, cin, k, k), 7.0, device=DEV)
    if shift_on_a:      # A = x (shifted), B = dy: D[m=ci][n=co]
        ops.wgrad_sps(xs, dys, n, P, taps, True, out, cin, cout, taps, cin * taps, 1)
    else:               # A = dy, B = x (shifted): D[m=co][n=ci]
        ops.wgrad_sps(dys, xs, n, P, taps, False, out, cout, cin, cin * taps, taps, 1)
    torch.cuda.synchronize()
    got = out.cpu()
    err = (got - want).abs().max().item() / want.abs().max().item()
    assert err <= 2e-3, err
    # accumulate mode adds onto what is there
    if shift_on_a:
        ops.wgrad_sps(xs, dys, n, P, taps, True, out, cin, cout, taps, cin * taps, 1, accumulate=True)
    else:
        ops.wgrad_sps(dys, xs, n, P, taps, False, out, cout, cin, cin * taps, taps, 1, accumulate=True)
    err2 = (out.cpu() - 2 * want).abs().max().item() / want.abs().max().item()
    assert err2 <= 4e-3, err2


def test_wgrad_bias_column():
    """A constant-one channel in B turns its column into the bias gradient (sum of dY rows)."""
    from vitcnn_b200 import ops
    n, P, cin, cout = 6, 11, 32, 96
    g = torch.Generator().manual_seed(3)
    x = emu.bf16(torch.randn(n, cin, P, P, generator=g))
    dy = emu.bf16(torch.randn(n, cout, P, P, generator=g))
    xa = torch.cat([x, torch.ones(n, 1, P, P)], 1)                       # channel 32 = ones
    xs = emu.pack_sps(xa, 6).to(torch.bfloat16).to(DEV)                  # 48 channels
    dys = emu.pack_sps(dy, 12).to(torch.bfloat16).to(DEV)
    dw = torch.zeros(cout, cin, device=DEV)
    db = torch.zeros(cout, device=DEV)
    ops.wgrad_sps(dys, xs, n, P, 1, False, dw, cout, cin, cin, 1, 0, bias_col=cin, out_bias=db)
    torch.cuda.synchronize()
    want_w = torch.einsum("bohw,bihw->oi", dy, x)
    want_b = dy.sum((0, 2, 3))
    assert (dw.cpu() - want_w).abs().max().item() <= 2e-3 * want_w.abs().max().item()
    assert (db.cpu() - want_b).abs().max().item() <= 2e-3 * max(1.0, want_b.abs().max().item())

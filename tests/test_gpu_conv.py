"""-m gpu: SPS packing and the tensor-core stem convolution vs a CPU emulation of the same
data layouts (tests/emu.py) and vs the SIMT twin kernel."""
import numpy as np
import pytest
import torch

from tests import emu

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _rand_conv(cin, cout, s_in, n_out, ns, taps, seed):
    from vitcnn_b200.model import pack_conv_weight
    g = torch.Generator().manual_seed(seed)
    k = 3 if taps == 9 else 1
    w = torch.randn(cout, cin, k, k, generator=g) / (cin * taps) ** 0.5
    scale = torch.zeros(n_out)
    bias = torch.zeros(n_out)
    scale[:cout] = 0.5 + torch.rand(cout, generator=g)
    bias[:cout] = 0.2 * torch.randn(cout, generator=g)
    return pack_conv_weight(w, s_in, n_out, ns), scale, bias


@pytest.mark.parametrize("C,P,n", [(144, 11, 5), (1, 11, 3), (64, 7, 9), (180, 15, 2), (20, 8, 4), (2, 5, 7)])
def test_pack_sps_bit_exact(C, P, n):
    from vitcnn_b200 import ops
    from vitcnn_b200.model import slices_for
    g = torch.Generator().manual_seed(C + P)
    x = torch.rand(n, C, P, P, generator=g)
    S = slices_for(C)
    want = emu.pack_sps(x, S).to(torch.bfloat16)
    got = ops.pack_sps(x.to(DEV), S).cpu()
    assert torch.equal(got.view(torch.int16), want.view(torch.int16))
    # the NHWC-memory / NCHW-view tensor test() builds (model_utils.py:1103-1106)
    xv = x.permute(0, 2, 3, 1).contiguous().to(DEV).permute(0, 3, 1, 2)
    got = ops.pack_sps(xv, S).cpu()
    assert torch.equal(got.view(torch.int16), want.view(torch.int16))


CONV_CASES = [  # cin, cout, n_out(padded), nsplit, taps, P, n
    (16, 16, 16, 1, 9, 5, 3),
    (16, 32, 32, 1, 9, 11, 2),
    (64, 32, 32, 1, 9, 11, 9),
    (128, 64, 64, 1, 9, 11, 7),
    (144, 128, 128, 2, 9, 11, 5),
    (1, 8, 16, 1, 9, 11, 4),
    (64, 32, 32, 1, 1, 7, 6),
    (180, 128, 128, 4, 9, 15, 3),
    (64, 128, 128, 2, 9, 7, 400),      # many tiles per CTA: persistent loop + phase wrap
]


@pytest.mark.parametrize("impl", [1, 0], ids=["simt", "tcgen05"])
@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_sps_vs_emulation(case, impl):
    from vitcnn_b200 import ops
    from vitcnn_b200.model import slices_for
    cin, cout, n_out, ns, taps, P, n = case
    s_in = slices_for(cin)
    g = torch.Generator().manual_seed(cin * 7 + P)
    x = torch.rand(n, cin, P, P, generator=g) - 0.3
    w, scale, bias = _rand_conv(cin, cout, s_in, n_out, ns, taps, seed=cin + cout)
    a = emu.pack_sps(x, s_in)
    want = emu.conv_sps(a, w, scale, bias, n, P, relu=True)
    got = ops.conv_sps(a.to(torch.bfloat16).to(DEV), w.to(DEV), scale.to(DEV), bias.to(DEV), n, P, relu=True,
                       impl=impl)
    torch.cuda.synchronize()
    got = got.float().cpu()
    H, M = emu.halo(P), 128 * emu.tiles(n, P)
    err = (got - want).abs().max().item()
    assert err <= 2e-2 * max(1.0, want.abs().max().item()), err
    # pad cells and tile-padding rows must be exact zeros: the next conv relies on them
    valid = torch.zeros(got.shape[1], dtype=torch.bool)
    valid[emu.row_index(n, P).reshape(-1)] = True
    assert got[:, H:H + M][:, ~valid[H:H + M]].abs().max().item() == 0.0

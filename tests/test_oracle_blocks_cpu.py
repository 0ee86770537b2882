"""The token stage of the model oracle (oracle/model_ref.py) against the reference's own ViT source:
outputs of the AST-extracted, unmodified timm classes (tests/golden/make_block_golden.py) committed as
tests/golden/block_golden.npz, plus a live bit-for-bit comparison when /root/reference is mounted."""
import os

import numpy as np
import pytest
import torch

from oracle.model_ref import ViTCNNRef
from tests.conftest import GOLDEN
from tests.golden import make_block_golden as G

TOKEN_KEYS = ("cls_token", "pos_embed", "blocks.", "norm.", "head.")


def _oracle_with(sd, P, K):
    torch.manual_seed(0)
    ref = ViTCNNRef(8, 1, patch_size=P, num_classes=K, dropout=0.0)
    missing, unexpected = ref.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all(not k.startswith(TOKEN_KEYS) for k in missing), missing     # only the CNN stem is left untouched
    return ref.eval()


def _stages(ref, x):
    with torch.no_grad():
        t0 = ref.embed_tokens(x)
        blk = ref.blocks[0]
        return dict(tokens=t0, block0=blk(t0), attn0=blk.attn(blk.norm1(t0)), mlp0=blk.mlp(blk.norm2(t0)),
                    logits=ref.forward_tokens(t0))


@pytest.mark.parametrize("case", [c[0] for c in G.CASES])
def test_token_stage_matches_the_extracted_reference_classes(case):
    g = np.load(os.path.join(GOLDEN, "block_golden.npz"))
    P, K, B, seed = [int(v) for v in g[f"{case}_cfg"]]
    sd = {k[len(case) + 4:]: torch.from_numpy(g[k]) for k in g.files if k.startswith(f"{case}_sd_")}
    ref = _oracle_with(sd, P, K)
    got = _stages(ref, torch.from_numpy(g[f"{case}_x"]))
    for k, v in got.items():
        want = torch.from_numpy(g[f"{case}_{k}"])
        assert v.shape == want.shape
        # same torch ops in the same order: equal up to the BLAS kernel the host CPU dispatches to
        assert (v - want).abs().max().item() <= 2e-6 * max(1.0, want.abs().max().item()), k


@pytest.mark.skipif(not G.available(), reason="/root/reference (vendored timm source) is not mounted")
def test_live_bit_for_bit_against_the_reference_source():
    Attention, Block, Mlp, TokenStage = G.extract()
    for name, P, K, B, seed in G.CASES:
        m = G.seeded_token_stage(TokenStage, P, K, seed)
        ref = _oracle_with(m.state_dict(), P, K)
        x = torch.randn(B, P * P, 32, generator=torch.Generator().manual_seed(7 + seed))
        with torch.no_grad():
            want = dict(tokens=m._pos_embed(x), logits=m(x))
            want["block0"] = m.blocks[0](want["tokens"])
        got = _stages(ref, x)
        for k, v in want.items():
            assert torch.equal(got[k], v), (name, k)
    # the oracle's block carries the reference's parameter names, so state_dicts are interchangeable
    blk = Block(dim=32, num_heads=4, qkv_bias=True)
    assert set(blk.state_dict()) == set(ViTCNNRef(8, 1, patch_size=5, num_classes=4).blocks[0].state_dict())


@pytest.mark.parametrize("case", [c[0] for c in G.CONV_CASES])
def test_stem_layer_matches_the_reference_conv_unit(case):
    """Every stem layer of the reconstruction is a conv_bn_relu; its arithmetic (eval with running statistics,
    training with batch statistics and the running-stat update) is pinned to FusAtNet.ConvUnit of the reference."""
    from oracle.model_ref import ConvBnRelu
    g = np.load(os.path.join(GOLDEN, "block_golden.npz"))
    cin, cout, P, B = [int(v) for v in g[f"{case}_cfg"]]
    layer = ConvBnRelu(cin, cout)
    layer.load_state_dict({k[len(case) + 4:]: torch.from_numpy(g[k]) for k in g.files if k.startswith(f"{case}_sd_")})
    x = torch.from_numpy(g[f"{case}_x"])
    with torch.no_grad():
        ev = layer.eval()(x)
    tr = layer.train()(x).detach()
    for got, key in ((ev, "eval"), (tr, "train"), (layer.bn.running_mean, "running_mean_after"),
                     (layer.bn.running_var, "running_var_after")):
        want = torch.from_numpy(g[f"{case}_{key}"])
        assert (got - want).abs().max().item() <= 2e-6 * max(1.0, want.abs().max().item()), key
    if G.available():
        live = G.seeded_conv_unit(G.extract_conv_unit(), cin, cout, cin)
        twin = ConvBnRelu(cin, cout)
        twin.load_state_dict(live.state_dict())
        with torch.no_grad():
            assert torch.equal(twin.eval()(x), live.eval()(x))


def test_tanh_gelu_deviation_is_bounded():
    """The CUDA token kernels evaluate GELU in its tanh form (csrc/vc_tokens.cuh gelu_tanh_approx) where the
    reference's nn.GELU is the exact erf form (layers/mlp.py:21).  This bounds what that substitution alone
    does to the logits, in fp32 on the oracle: far inside the 2e-2 parity budget."""
    import torch.nn.functional as F
    from oracle import model_ref as M
    torch.manual_seed(3)
    ref = ViTCNNRef(32, 1, patch_size=9, num_classes=8, dropout=0.0).eval()
    with torch.no_grad():
        for blk in ref.blocks:                       # realistic (trained-scale) MLP pre-activations
            blk.mlp.fc1.weight.mul_(20.0)
            blk.attn.qkv.weight.mul_(8.0)
    g = torch.Generator().manual_seed(4)
    hsi, lid = torch.rand(16, 32, 9, 9, generator=g), torch.rand(16, 1, 9, 9, generator=g)
    with torch.no_grad():
        exact = ref(hsi, lid)
        orig = F.gelu
        try:
            M.F.gelu = lambda x: orig(x, approximate="tanh")
            approx = ref(hsi, lid)
        finally:
            M.F.gelu = orig
    assert not torch.equal(exact, approx)
    dev = (approx - exact).abs().max().item() / exact.abs().max().item()
    assert dev <= 2e-3, dev

"""not-gpu: the Python packing (BN fold, weight layout, token blob, SPS geometry) reproduces
the fp32 oracle when pushed through a CPU emulation of the kernels' data layouts, and the
C-ABI library loads and exports every symbol include/vitcnn.h declares."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import vitcnn_b200
from oracle.model_ref import ViTCNNRef, randomize_bn_stats
from tests import emu
from vitcnn_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _pair(C1, C2, P, K, seed=0):
    torch.manual_seed(seed)
    ref = ViTCNNRef(C1, C2, patch_size=P, num_classes=K)
    randomize_bn_stats(ref, seed=1)
    with torch.no_grad():   # non-trivial LN / bias parameters so every blob field matters
        for p in ref.parameters():
            if p.dim() == 1:
                p.add_(0.05 * torch.randn_like(p))
        ref.cls_token.normal_(std=0.02)
    ref.eval()
    ours = vitcnn_b200.ViTCNN(C1, C2, patch_size=P, num_classes=K)
    missing = ours.load_state_dict(ref.state_dict(), strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    return ref, ours.eval()


def test_state_dict_keys_match_oracle():
    ref, ours = _pair(16, 1, 5, 4)
    assert list(ref.state_dict().keys()) == list(ours.state_dict().keys())
    for k, v in ref.state_dict().items():
        assert ours.state_dict()[k].shape == v.shape


@pytest.mark.parametrize("cfg", [(16, 1, 5, 4, 3), (64, 2, 7, 12, 2), (20, 1, 8, 5, 2), (144, 1, 11, 16, 2)])
def test_emulated_kernels_match_oracle(cfg):
    C1, C2, P, K, B = cfg
    ref, ours = _pair(C1, C2, P, K)
    g = torch.Generator().manual_seed(5)
    hsi, lid = torch.rand(B, C1, P, P, generator=g), torch.rand(B, C2, P, P, generator=g)
    with torch.no_grad():
        want = ref(hsi, lid)
        got = emu.model_forward(ours, emu.py_tparams_layout(P, K), hsi, lid)
    scale = want.abs().max().item()
    assert (got - want).abs().max().item() <= 2e-2 * scale, ((got - want).abs().max().item(), scale)


def test_library_exports_every_declared_symbol():
    _lib.build()
    header = open(os.path.join(ROOT, "include", "vitcnn.h")).read()
    declared = set(re.findall(r"\b(vc_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.EXPORTED), declared ^ set(_lib.EXPORTED)
    L = ctypes.CDLL(_lib.SO_PATH)
    for name in declared:
        assert hasattr(L, name), name
    assert _lib.lib().vc_abi_version() == 4


def test_layout_helpers_match_c():
    _lib.build()
    for P, K in [(5, 4), (7, 12), (11, 16), (15, 8)]:
        c = _lib.tparams_layout(P, K)
        py = emu.py_tparams_layout(P, K)
        assert c["total"] == py["total"] and c["pos"] == py["pos"] and c["layers"] == py["layers"]
        for n in (1, 7, 64, 1000):
            assert _lib.lib().vc_sps_rows(n, P) == emu.rows(n, P)


def test_cpu_forward_refuses():
    _, ours = _pair(16, 1, 5, 4)
    with pytest.raises(RuntimeError, match="CUDA"):
        ours(torch.rand(1, 16, 5, 5), torch.rand(1, 1, 5, 5))


def test_get_model_contract():
    hp = dict(n_classes=16, n_bands=(144, 1), ignored_labels=[0], dataset="Houston2013")
    model, opt, crit, out = vitcnn_b200.get_model("ViT-CNN", **hp)
    assert isinstance(model, vitcnn_b200.ViTCNN) and isinstance(opt, torch.optim.Adam)
    assert out["patch_size"] == 11 and out["center_pixel"] is True and out["batch_size"] == 64
    assert out["epoch"] == 128 and out["applyPCA"] is False and out["supervision"] == "full"
    assert crit.weight[0] == 0 and crit.weight[1:].eq(1).all()
    assert sum(p.numel() for p in model.parameters()) == 296800
    with pytest.raises(KeyError, match="model is unknown"):
        vitcnn_b200.get_model("nope", **hp)

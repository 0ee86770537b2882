"""-m gpu: the CUDA module vs the fp32 oracle with the same state_dict:
logits within 2e-2 relative (bf16 compute, fp32 accumulate), argmax agreement >= 99.9 % on
structured data, drop-in behaviours the reference's loops rely on."""
import copy

import numpy as np
import pytest
import torch

from oracle import data_ref as R
from oracle.model_ref import ViTCNNRef, randomize_bn_stats

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
REL_TOL = 2e-2   # north star: logits within 2e-2 relative of the fp32 reference


def make_pair(C1, C2, P, K, seed=0):
    import vitcnn_b200
    torch.manual_seed(seed)
    ref = ViTCNNRef(C1, C2, patch_size=P, num_classes=K)
    randomize_bn_stats(ref, seed=1)
    with torch.no_grad():
        for p in ref.parameters():
            if p.dim() == 1:
                p.add_(0.05 * torch.randn_like(p))
    ref.eval()
    ours = vitcnn_b200.ViTCNN(C1, C2, patch_size=P, num_classes=K)
    ours.load_state_dict(ref.state_dict())
    return ref, ours.to(DEV).eval()


def rel_err(got, want):
    return (got - want).abs().max().item() / max(want.abs().max().item(), 1e-6)


@pytest.mark.parametrize("cfg", [(16, 1, 5, 4, 9), (64, 2, 7, 12, 33), (144, 1, 9, 16, 17), (144, 1, 11, 16, 64),
                                 (180, 1, 11, 8, 10), (64, 2, 11, 12, 5), (144, 1, 15, 16, 6), (20, 1, 8, 5, 4)])
def test_logits_vs_oracle(cfg):
    C1, C2, P, K, B = cfg
    ref, ours = make_pair(C1, C2, P, K)
    g = torch.Generator().manual_seed(5)
    hsi, lid = torch.rand(B, C1, P, P, generator=g), torch.rand(B, C2, P, P, generator=g)
    with torch.no_grad():
        want = ref(hsi, lid)
        got = ours(hsi.to(DEV), lid.to(DEV)).cpu()
    assert got.dtype == torch.float32 and got.shape == want.shape
    assert rel_err(got, want) <= REL_TOL, rel_err(got, want)


def test_golden_logits(model_golden):
    """Committed oracle outputs (weights in the fixture): guards the oracle and the kernels
    against drifting together."""
    import vitcnn_b200
    g = model_golden
    for name in ("small", "muufl7"):
        C1, C2, P, K, B = [int(v) for v in g[f"{name}_cfg"]]
        ours = vitcnn_b200.ViTCNN(C1, C2, patch_size=P, num_classes=K)
        sd = {k[len(name) + 4:]: torch.from_numpy(g[k]) for k in g.files if k.startswith(f"{name}_sd_")}
        ours.load_state_dict(sd)
        ours = ours.to(DEV).eval()
        with torch.no_grad():
            got = ours(torch.from_numpy(g[f"{name}_hsi"]).to(DEV), torch.from_numpy(g[f"{name}_lidar"]).to(DEV)).cpu()
        assert rel_err(got, torch.from_numpy(g[f"{name}_logits"])) <= REL_TOL


def test_strided_inputs_and_short_batch():
    """test() feeds an NCHW view over NHWC memory and a short last batch
    (model_utils.py:1103-1112); values must not depend on strides or batch size."""
    ref, ours = make_pair(144, 1, 11, 16)
    g = torch.Generator().manual_seed(9)
    hsi, lid = torch.rand(37, 144, 11, 11, generator=g), torch.rand(37, 1, 11, 11, generator=g)
    with torch.no_grad():
        a = ours(hsi.to(DEV), lid.to(DEV))
        hv = hsi.permute(0, 2, 3, 1).contiguous().to(DEV).permute(0, 3, 1, 2)
        lv = lid.permute(0, 2, 3, 1).contiguous().to(DEV).permute(0, 3, 1, 2)
        b = ours(hv, lv)
        c = torch.cat([ours(hsi[:20].to(DEV), lid[:20].to(DEV)), ours(hsi[20:].to(DEV), lid[20:].to(DEV))])
        e = ours(hsi[:0].to(DEV), lid[:0].to(DEV))
    assert torch.equal(a, b) and torch.equal(a, c) and e.shape == (0, 16)


def test_argmax_agreement_on_structured_scene():
    """>= 99.9 % argmax agreement (north star) on structured synthetic data after a brief
    oracle training run, so logit margins are meaningful."""
    import vitcnn_b200
    C1, C2, P, K = 32, 1, 7, 6
    img1, img2, gt = R.synthetic_scene(48, 64, C1, C2, K, seed=2)
    torch.manual_seed(0)
    ref = ViTCNNRef(C1, C2, patch_size=P, num_classes=K, dropout=0.0)
    idx = R.train_indices(gt, [0], P)
    rng = np.random.default_rng(0)
    opt = torch.optim.Adam(ref.parameters(), lr=2e-3)
    w = torch.ones(K)
    w[0] = 0
    ref.train()
    for _ in range(60):
        sel = idx[rng.choice(len(idx), 64)]
        h, l, y = R.gather_centers(img1, img2, gt, sel, P)
        loss = torch.nn.functional.cross_entropy(ref(torch.from_numpy(h), torch.from_numpy(l)), torch.from_numpy(y), w)
        opt.zero_grad()
        loss.backward()
        opt.step()
    ref.eval()
    ours = vitcnn_b200.ViTCNN(C1, C2, patch_size=P, num_classes=K)
    ours.load_state_dict(ref.state_dict())
    ours = ours.to(DEV).eval()
    corners = R.sliding_window_corners(gt.shape, 1, (P, P))
    h, l = R.gather_corners(img1, img2, corners, P)
    with torch.no_grad():
        want = ref(torch.from_numpy(h), torch.from_numpy(l))
        got = ours(torch.from_numpy(h).to(DEV), torch.from_numpy(l).to(DEV)).cpu()
    assert rel_err(got, want) <= REL_TOL
    agree = (got.argmax(1) == want.argmax(1)).float().mean().item()
    assert agree >= 0.999, agree


def test_module_contract():
    """What train()/main.py do with the module: deepcopy of state_dict, reload, class name."""
    ref, ours = make_pair(16, 1, 5, 4)
    sd = copy.deepcopy(ours.state_dict())
    g = torch.Generator().manual_seed(1)
    hsi, lid = torch.rand(3, 16, 5, 5, generator=g).to(DEV), torch.rand(3, 1, 5, 5, generator=g).to(DEV)
    with torch.no_grad():
        a = ours(hsi, lid)
        with torch.no_grad():
            for p in ours.parameters():
                p.mul_(1.5)
        b = ours(hsi, lid)            # repacked after the in-place update
        ours.load_state_dict(sd)
        c = ours(hsi, lid)
    assert not torch.equal(a, b) and torch.equal(a, c)
    assert type(ours).__name__ == "ViTCNN"

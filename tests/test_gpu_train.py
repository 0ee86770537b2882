"""-m gpu: training kernels vs torch autograd on the fp32 oracle.

Tolerances: activations / gradients pass through bf16 operands with fp32 accumulation, so
per-tensor gradients are compared as max-abs error relative to the tensor's max-abs value
(GRAD_TOL) plus a cosine-similarity floor; integer / fp32-only pieces (loss, Adam, BatchNorm
statistics) are compared tightly."""
import ctypes

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import data_ref as R
from oracle.model_ref import ViTCNNRef, randomize_bn_stats
from tests import emu

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GRAD_TOL = 3e-2      # floor; the yardstick is torch's own bf16 autocast of the oracle (see _autocast_grads)
COS_MIN = 0.999


def _stream():
    return torch.cuda.current_stream().cuda_stream


@pytest.mark.parametrize("C,P,n", [(32, 11, 6), (8, 5, 9), (128, 7, 20)])
def test_bn_forward_backward(C, P, n):
    from vitcnn_b200 import _lib
    L = _lib.lib()
    S = (C + 15) // 16 * 2
    g = torch.Generator().manual_seed(C + P)
    y = emu.bf16(torch.randn(n, C, P, P, generator=g) * 1.5 + 0.3)
    dz = emu.bf16(torch.randn(n, C, P, P, generator=g))
    gamma = 0.5 + torch.rand(C, generator=g)
    beta = 0.3 * torch.randn(C, generator=g)
    bn = torch.nn.BatchNorm2d(C)
    with torch.no_grad():
        bn.weight.copy_(gamma)
        bn.bias.copy_(beta)
    bn.train()
    yr = y.clone().requires_grad_(True)
    zr = F.relu(bn(yr))
    zr.backward(dz)
    ys = emu.pack_sps(y, S).to(torch.bfloat16).to(DEV)
    dzs = emu.pack_sps(dz, S).to(torch.bfloat16).to(DEV)
    zs = torch.empty_like(ys)
    rm, rv = torch.zeros(C, device=DEV), torch.ones(C, device=DEV)
    nbt = torch.zeros((), dtype=torch.int64, device=DEV)
    sums = torch.zeros(2 * S * 8, dtype=torch.float64, device=DEV)
    scale, shift, mean, rstd = (torch.empty(S * 8, device=DEV) for _ in range(4))
    gd, bd = gamma.to(DEV), beta.to(DEV)
    _lib.check(L.vc_bn_forward(ys.data_ptr(), zs.data_ptr(), S, C, n, P, gd.data_ptr(), bd.data_ptr(), 1e-5, 0.1,
                               rm.data_ptr(), rv.data_ptr(), nbt.data_ptr(), sums.data_ptr(), scale.data_ptr(),
                               shift.data_ptr(), mean.data_ptr(), rstd.data_ptr(), 1, _stream()), "bn fwd")
    got_z = emu.unpack_sps(zs.cpu(), n, P, C)
    assert (got_z - zr.detach()).abs().max().item() <= 2e-2 * zr.abs().max().item()
    assert torch.allclose(rm.cpu(), bn.running_mean, atol=1e-4) and torch.allclose(rv.cpu(), bn.running_var, atol=1e-4)
    assert int(nbt.item()) == 1 and float(sums.abs().max().item()) == 0.0
    # pad cells stay zero
    valid = torch.zeros(zs.shape[1], dtype=torch.bool)
    valid[emu.row_index(n, P).reshape(-1)] = True
    assert zs.cpu().float()[:, ~valid].abs().max().item() == 0.0
    dys = torch.empty_like(ys)
    dgam, dbet, dbias = (torch.full((C,), 3.0, device=DEV) for _ in range(3))
    _lib.check(L.vc_bn_backward(dzs.data_ptr(), ys.data_ptr(), dys.data_ptr(), S, C, n, P, scale.data_ptr(),
                                shift.data_ptr(), mean.data_ptr(), rstd.data_ptr(), 1, sums.data_ptr(), dgam.data_ptr(),
                                dbet.data_ptr(), dbias.data_ptr(), _stream()), "bn bwd")
    got_dy = emu.unpack_sps(dys.cpu(), n, P, C)
    assert (got_dy - yr.grad).abs().max().item() <= 2e-2 * yr.grad.abs().max().item()
    assert (dgam.cpu() - bn.weight.grad).abs().max().item() <= 1e-2 * bn.weight.grad.abs().max().item()
    assert (dbet.cpu() - bn.bias.grad).abs().max().item() <= 1e-2 * bn.bias.grad.abs().max().item()
    assert dbias.abs().max().item() == 0.0 and float(sums.abs().max().item()) == 0.0


def test_ce_loss_and_grad():
    from vitcnn_b200.train import ce_loss
    g = torch.Generator().manual_seed(0)
    n, K = 777, 16
    logits = (3 * torch.randn(n, K, generator=g)).requires_grad_(True)
    labels = torch.randint(0, K, (n,), generator=g)
    labels[5] = -100                                      # ignore_index
    w = torch.ones(K)
    w[0] = 0.0
    w[3] = 2.5
    want = F.cross_entropy(logits, labels, weight=w)
    want.backward()
    loss, d = ce_loss(logits.detach().to(DEV), labels.to(DEV), w.to(DEV))
    assert abs(loss[0].item() - want.item()) <= 1e-5 * max(1.0, abs(want.item()))
    assert (d.cpu() - logits.grad).abs().max().item() <= 1e-6
    loss2, d2 = ce_loss(logits.detach().to(DEV), labels.to(DEV), None, grad_scale=0.5)
    l2 = F.cross_entropy(logits.detach().requires_grad_(True), labels)
    assert abs(loss2[0].item() - l2.item()) <= 1e-5 * max(1.0, abs(l2.item()))


def test_adam_matches_torch():
    from vitcnn_b200.train import adam_step
    g = torch.Generator().manual_seed(1)
    p0 = torch.randn(10007, generator=g)
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([ref], lr=1e-3)
    p, m, v = p0.clone().to(DEV), torch.zeros(10007, device=DEV), torch.zeros(10007, device=DEV)
    for t in range(1, 6):
        gr = torch.randn(10007, generator=g) * (10.0 ** (t - 3))
        ref.grad = gr.clone()
        opt.step()
        adam_step(p, (2 * gr).to(DEV), m, v, t, lr=1e-3, grad_scale=0.5)
    assert (p.cpu() - ref.detach()).abs().max().item() <= 2e-6


@pytest.mark.parametrize("case", [(128, 64, 9, 11, 4), (32, 16, 9, 7, 5), (64, 32, 1, 11, 3)])
def test_dgrad_conv_via_transposed_pack(case):
    """Data gradient of a conv = the same tensor-core conv with the transposed / flipped operand."""
    from vitcnn_b200 import _lib, ops
    from vitcnn_b200.model import choose_nsplit, pack_conv_weight, slices_for
    cin, cout, taps, P, n = case
    k = 3 if taps == 9 else 1
    g = torch.Generator().manual_seed(cin + cout)
    w = torch.randn(cout, cin, k, k, generator=g) / (cin * taps) ** 0.5
    dy = emu.bf16(torch.randn(n, cout, P, P, generator=g))
    x = torch.zeros(n, cin, P, P, requires_grad=True)
    F.conv2d(x, emu.bf16(w), padding=k // 2).backward(dy)
    want = x.grad
    s_in, n_out = slices_for(cout), slices_for(cin) * 8
    ns = choose_nsplit(s_in, n_out, taps)
    wd = torch.empty(ns, taps, s_in, n_out // ns, 8, dtype=torch.bfloat16, device=DEV)
    wdev = w.to(DEV).contiguous()
    _lib.check(_lib.lib().vc_pack_conv_weight(wdev.data_ptr(), cout, cin, taps, 1, s_in, n_out, ns, wd.data_ptr(), _stream()),
               "pack")
    # forward operand from the same kernel must equal the Python packing
    ns_f = choose_nsplit(slices_for(cin), slices_for(cout) * 8, taps)
    wf = torch.empty(ns_f, taps, slices_for(cin), slices_for(cout) * 8 // ns_f, 8, dtype=torch.bfloat16, device=DEV)
    _lib.check(_lib.lib().vc_pack_conv_weight(wdev.data_ptr(), cout, cin, taps, 0, slices_for(cin), slices_for(cout) * 8, ns_f,
                                              wf.data_ptr(), _stream()), "pack")
    assert torch.equal(wf.cpu().view(torch.int16),
                       pack_conv_weight(w, slices_for(cin), slices_for(cout) * 8, ns_f).view(torch.int16))
    dys = emu.pack_sps(dy, s_in).to(torch.bfloat16).to(DEV)
    ones, zeros = torch.ones(n_out, device=DEV), torch.zeros(n_out, device=DEV)
    got = ops.conv_sps(dys, wd, ones, zeros, n, P, relu=False)
    got = emu.unpack_sps(got.cpu(), n, P, cin)
    assert (got - want).abs().max().item() <= 2e-2 * want.abs().max().item()


def _pair(C1, C2, P, K, seed=0):
    import vitcnn_b200
    torch.manual_seed(seed)
    ref = ViTCNNRef(C1, C2, patch_size=P, num_classes=K, dropout=0.0)
    randomize_bn_stats(ref, seed=1)
    with torch.no_grad():
        for p in ref.parameters():
            if p.dim() == 1:
                p.add_(0.05 * torch.randn_like(p))
        ref.cls_token.normal_(std=0.02)
        for blk in ref.blocks:                     # default init (std 0.02) makes attention nearly uniform:
            blk.attn.qkv.weight.mul_(8.0)          # sharpen it so the softmax backward is exercised
    ours = vitcnn_b200.ViTCNN(C1, C2, patch_size=P, num_classes=K, dropout=0.0)
    ours.load_state_dict(ref.state_dict())
    return ref.train(), ours.to(DEV).train()


def _grad_report(ref, ours):
    rows = []
    named = dict(ours.named_parameters())
    for k, p in ref.named_parameters():
        g, w = named[k].grad, p.grad
        assert g is not None, k
        g = g.detach().cpu()
        scale = w.abs().max().item()
        err = (g - w).abs().max().item() / max(scale, 1e-12)
        cos = F.cosine_similarity(g.reshape(1, -1), w.reshape(1, -1)).item() if scale > 0 else 1.0
        rows.append((k, err, cos, scale))
    return rows


def _autocast_grads(ref, hsi, lid, y, w):
    """Noise floor of bf16 compute for this batch: the oracle itself under torch.autocast(bf16)
    against its fp32 gradients.  With BatchNorm's mean subtraction the conv-stem gradients are
    small differences of large sums, so bf16 rounding of the activations shows up as 10-15 %
    of max-abs there on random data, for ANY bf16 implementation."""
    import copy
    low = copy.deepcopy(ref)
    low.zero_grad()
    with torch.autocast("cpu", dtype=torch.bfloat16):
        out = low(hsi, lid)
    F.cross_entropy(out.float(), y, weight=w).backward()
    floor = {}
    for (k, p), (_, q) in zip(ref.named_parameters(), low.named_parameters()):
        s = p.grad.abs().max().item()
        floor[k] = ((q.grad - p.grad).abs().max().item() / max(s, 1e-12),
                    F.cosine_similarity(q.grad.reshape(1, -1), p.grad.reshape(1, -1)).item())
    return floor


@pytest.mark.parametrize("cfg", [(16, 1, 5, 4, 6), (64, 2, 7, 12, 9), (144, 1, 11, 16, 8), (180, 1, 11, 8, 5),
                                 (24, 1, 9, 6, 7), (16, 1, 13, 4, 4), (20, 2, 15, 5, 3)])
def test_model_gradients_vs_oracle_autograd(cfg):
    C1, C2, P, K, B = cfg
    ref, ours = _pair(C1, C2, P, K)
    g = torch.Generator().manual_seed(5)
    hsi, lid = torch.rand(B, C1, P, P, generator=g), torch.rand(B, C2, P, P, generator=g)
    y = torch.randint(1, K, (B,), generator=g)
    w = torch.ones(K)
    w[0] = 0
    lr_ = F.cross_entropy(ref(hsi, lid), y, weight=w)
    lr_.backward()
    out = ours(hsi.to(DEV), lid.to(DEV))
    assert out.requires_grad and out.dtype == torch.float32
    lo = F.cross_entropy(out, y.to(DEV), weight=w.to(DEV))
    lo.backward()
    assert abs(lo.item() - lr_.item()) <= 2e-2 * max(1.0, abs(lr_.item()))
    rows = _grad_report(ref, ours)
    floor = _autocast_grads(ref, hsi, lid, y, w)
    rows = [r for r in rows if not r[0].endswith("conv.bias") and r[3] > 1e-7]
    bad = [(k, round(e, 4), round(c, 5), floor[k]) for k, e, c, s in rows
           if e > max(GRAD_TOL, 4.0 * floor[k][0]) or c < min(0.98, floor[k][1] - 0.02)]
    assert not bad, bad
    # no systematic excess over the bf16 floor, and the whole gradient points the same way
    ratios = sorted(e / max(floor[k][0], 1e-3) for k, e, c, s in rows)
    assert ratios[len(ratios) // 2] <= 1.5, ratios
    named = dict(ours.named_parameters())
    flat_o = torch.cat([named[k].grad.detach().cpu().reshape(-1) for k, _ in ref.named_parameters()])
    flat_r = torch.cat([p.grad.reshape(-1) for _, p in ref.named_parameters()])
    assert F.cosine_similarity(flat_o.view(1, -1), flat_r.view(1, -1)).item() >= COS_MIN
    # conv biases feed a BatchNorm: their gradient is analytically zero
    for k, e, c, s in rows:
        if k.endswith("conv.bias"):
            assert dict(ours.named_parameters())[k].grad.abs().max().item() <= 1e-6
    # running statistics follow nn.BatchNorm2d
    rb, ob = dict(ref.named_buffers()), dict(ours.named_buffers())
    for k in rb:
        if k.endswith("num_batches_tracked"):
            assert int(ob[k].item()) == int(rb[k].item())
        else:
            assert (ob[k].cpu() - rb[k]).abs().max().item() <= 2e-2 * max(1.0, rb[k].abs().max().item()), k


def test_reference_style_loop_and_trainer_converge():
    """A few epochs of the reference-style loop (torch CE + optim.Adam on our module) and of the
    fused Trainer on a small structured scene: loss drops like the fp32 oracle's."""
    import vitcnn_b200
    from vitcnn_b200.train import Trainer
    C1, C2, P, K = 32, 1, 7, 6
    img1, img2, gt = R.synthetic_scene(40, 56, C1, C2, K, seed=3)
    idx = R.train_indices(gt, [0], P)
    rng = np.random.default_rng(0)
    w = torch.ones(K)
    w[0] = 0
    torch.manual_seed(0)
    ref = ViTCNNRef(C1, C2, patch_size=P, num_classes=K, dropout=0.0).train()
    ours = vitcnn_b200.ViTCNN(C1, C2, patch_size=P, num_classes=K, dropout=0.0)
    ours.load_state_dict(ref.state_dict())
    ours = ours.to(DEV).train()
    fused = vitcnn_b200.ViTCNN(C1, C2, patch_size=P, num_classes=K, dropout=0.0)
    fused.load_state_dict(ref.state_dict())
    fused = fused.to(DEV)
    tr = Trainer(fused, lr=2e-3, weights=w)
    t1, t2, tg = torch.from_numpy(img1).to(DEV), torch.from_numpy(img2).to(DEV), torch.from_numpy(gt).to(DEV)
    opt_r = torch.optim.Adam(ref.parameters(), lr=2e-3)
    opt_o = torch.optim.Adam(ours.parameters(), lr=2e-3)
    lr_hist, lo_hist, lf_hist = [], [], []
    for it in range(40):
        sel = idx[rng.choice(len(idx), 64)]
        h, l, y = R.gather_centers(img1, img2, gt, sel, P)
        loss = F.cross_entropy(ref(torch.from_numpy(h), torch.from_numpy(l)), torch.from_numpy(y), w)
        opt_r.zero_grad(); loss.backward(); opt_r.step()
        lr_hist.append(loss.item())
        lo = F.cross_entropy(ours(torch.from_numpy(h).to(DEV), torch.from_numpy(l).to(DEV)), torch.from_numpy(y).to(DEV),
                             w.to(DEV))
        opt_o.zero_grad(); lo.backward(); opt_o.step()
        lo_hist.append(lo.item())
        lf = tr.step(t1, t2, tg, torch.from_numpy(sel.astype(np.int32)).to(DEV))
        lf_hist.append(lf[0].item())
    first, last = np.mean(lr_hist[:5]), np.mean(lr_hist[-5:])
    assert last < 0.6 * first                                         # the oracle itself learns
    assert abs(np.mean(lo_hist[-5:]) - last) <= 0.25 * first, (lo_hist[-5:], lr_hist[-5:])
    assert abs(np.mean(lf_hist[-5:]) - last) <= 0.25 * first, (lf_hist[-5:], lr_hist[-5:])
    assert abs(lo_hist[0] - lr_hist[0]) <= 2e-2 * lr_hist[0] and abs(lf_hist[0] - lr_hist[0]) <= 2e-2 * lr_hist[0]
    # eval after training uses the updated weights and running statistics
    fused.eval(); ref.eval()
    with torch.no_grad():
        a = fused(torch.from_numpy(h).to(DEV), torch.from_numpy(l).to(DEV)).cpu()
    assert torch.isfinite(a).all()


def test_trainer_cuda_graph_matches_eager():
    """The captured training step (every kernel + Adam with device-side step counter) replays to
    the same parameters as eager launches (up to the order of fp32 atomics in the LayerNorm /
    pos-embed gradient sums)."""
    import vitcnn_b200
    from vitcnn_b200.train import Trainer
    C1, C2, P, K = 32, 1, 7, 6
    img1, img2, gt = R.synthetic_scene(40, 56, C1, C2, K, seed=3)
    idx = R.train_indices(gt, [0], P)
    rng = np.random.default_rng(1)
    t1, t2, tg = torch.from_numpy(img1).to(DEV), torch.from_numpy(img2).to(DEV), torch.from_numpy(gt).to(DEV)
    w = torch.ones(K)
    w[0] = 0
    torch.manual_seed(0)
    a = vitcnn_b200.ViTCNN(C1, C2, patch_size=P, num_classes=K, dropout=0.0)
    b = vitcnn_b200.ViTCNN(C1, C2, patch_size=P, num_classes=K, dropout=0.0)
    b.load_state_dict(a.state_dict())
    ta = Trainer(a.to(DEV), lr=1e-3, weights=w)
    tb = Trainer(b.to(DEV), lr=1e-3, weights=w, use_graph=True, graph_warmup=2)
    la, lb = [], []
    for it in range(8):
        xy = torch.from_numpy(idx[rng.choice(len(idx), 48)].astype(np.int32)).to(DEV)
        la.append(ta.step(t1, t2, tg, xy)[0].item())
        lb.append(tb.step(t1, t2, tg, xy)[0].item())
        if it == 4:
            ta.set_lr(5e-4)
            tb.set_lr(5e-4)
    assert len(tb._graphs) == 1 and ta.t == 8 and tb.t == 8
    assert np.allclose(la, lb, rtol=2e-3, atol=2e-4), (la, lb)
    fa, fb = ta.state.flat, tb.state.flat
    # Adam moves a parameter by up to lr per step whatever the gradient's size, so atomics-order noise on
    # near-zero gradients shows as differences of a few lr; the bulk of the parameters agree closely
    assert (fa - fb).abs().max().item() <= 8e-3 and (fa - fb).abs().mean().item() <= 1e-4


@pytest.mark.parametrize("p_drop", [0.0, 0.3])
def test_dropout_masks_consistent_between_forward_and_backward(p_drop):
    """Dropout masks are a stateless hash of (seed, sample, site, element) evaluated by the forward,
    the recompute and the backward: (1) the same seed gives the same logits, another seed different
    ones; (2) with the seed pinned, the directional derivative of the loss along our gradient
    matches a central finite difference (a mask mismatch anywhere would break this; p = 0 is the
    control for the bf16 noise of the method)."""
    import vitcnn_b200
    from vitcnn_b200.train import train_state
    C1, C2, P, K, B = 16, 1, 7, 5, 48
    torch.manual_seed(0)
    net = vitcnn_b200.ViTCNN(C1, C2, patch_size=P, num_classes=K, dropout=p_drop)
    with torch.no_grad():
        for blk in net.blocks:
            blk.attn.qkv.weight.mul_(8.0)
        net.cls_token.normal_(std=0.02)
    net = net.to(DEV).train()
    g = torch.Generator().manual_seed(2)
    hsi, lid = torch.rand(B, C1, P, P, generator=g).to(DEV), torch.rand(B, C2, P, P, generator=g).to(DEV)
    y = torch.randint(1, K, (B,), generator=g).to(DEV)
    st = train_state(net)
    seed0 = 12345

    def loss_at():
        st.drop_seed.fill_(seed0)
        with torch.no_grad():
            return F.cross_entropy(net(hsi, lid).double(), y).item()

    st.drop_seed.fill_(seed0)
    out1 = net(hsi, lid)
    F.cross_entropy(out1, y).backward()
    grad = torch.cat([p.grad.reshape(-1) for p in st.params]).double()
    st.drop_seed.fill_(seed0)
    with torch.no_grad():
        out2 = net(hsi, lid)
        st.drop_seed.fill_(seed0 + 1)
        out3 = net(hsi, lid)
    assert torch.equal(out1.detach(), out2)
    assert (not torch.equal(out2, out3)) == (p_drop > 0)
    gn = grad.norm().item()
    dirs = [(p_.grad / gn).clone() for p_ in st.params]
    eps = 0.04 / gn
    with torch.no_grad():
        def shift(scale):
            for p_, d_ in zip(st.params, dirs):
                p_.add_(scale * d_)
        shift(eps)
        lp = loss_at()
        shift(-2 * eps)
        lm = loss_at()
        shift(eps)
    fd = (lp - lm) / (2 * eps)
    assert abs(fd - gn) <= 0.15 * gn, (fd, gn, p_drop)

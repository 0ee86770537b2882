"""-m gpu: the reference's OWN loops drive the CUDA module.

``model_utils.test()`` (model_utils.py:1067-1132), ``val()`` (:1135-1158) and ``train()`` (:854-1045) are
imported unmodified -- from /root/reference in the dev container, from the git-ignored ``baseline/_ref/``
(tools/install_ref.sh) on the GPU box -- with the stub recipe of SURVEY App. B (oracle/ref_import.py), and
run with ``vitcnn_b200.ViTCNN`` as ``net``.  Their results are compared with this package's mirrors
(``vitcnn_b200.test / val / train``, the drop-in replacements INTEGRATION.md names)."""
import copy
import os

import numpy as np
import pytest
import torch

from oracle import data_ref as R
from oracle import ref_import
from oracle.model_ref import ViTCNNRef, randomize_bn_stats

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not ref_import.available(), reason="no reference tree (run tools/install_ref.sh)")]
DEV = "cuda:0"


@pytest.fixture(scope="module")
def ref_mods():
    return ref_import.import_reference()


def _net(C1, C2, P, K, seed=0, dropout=0.01):
    import vitcnn_b200
    torch.manual_seed(seed)
    ref = ViTCNNRef(C1, C2, patch_size=P, num_classes=K, dropout=dropout)
    randomize_bn_stats(ref, seed=1)
    net = vitcnn_b200.ViTCNN(C1, C2, patch_size=P, num_classes=K, dropout=dropout)
    net.load_state_dict(ref.state_dict())
    return net.to(DEV)


@pytest.mark.parametrize("cfg", [(27, 40, 144, 1, 11, 16, 1, 64), (30, 26, 64, 2, 7, 12, 2, 50)])
def test_reference_test_loop_equals_the_scene_path_bit_for_bit(ref_mods, cfg):
    """model_utils.test() -- sliding_window -> grouper -> np.copy -> NCHW view over NHWC memory -> net() ->
    probs[x + w//2, y + h//2] += out -- on our module, against vitcnn_b200.test() (one vc_scene_infer call)."""
    import vitcnn_b200
    _, _, ref_mu = ref_mods
    H, W, C1, C2, P, K, stride, bs = cfg
    net = _net(C1, C2, P, K).eval()
    img1, img2, _ = R.synthetic_scene(H, W, C1, C2, K, seed=5)
    hp = dict(patch_size=P, center_pixel=True, batch_size=bs, device=torch.device(DEV), n_classes=K, applyPCA=False,
              test_stride=stride)
    want = ref_mu.test(0, net, img1, img2, hp)
    got = vitcnn_b200.test(0, net, img1, img2, hp)
    assert want.dtype == got.dtype == np.float64 and want.shape == got.shape == (H, W, K)
    assert (want != 0).any()
    assert np.array_equal(got, want)


def _datasets(ref_ds, img1, img2, gt, P, device, **flags):
    import vitcnn_b200
    hp = dict(dataset="synthetic", patch_size=P, ignored_labels=[0], flip_augmentation=False, radiation_augmentation=False,
              mixture_augmentation=False, center_pixel=True, supervision="full", applyPCA=False)
    hp.update(flags)
    np.random.seed(11)
    theirs = ref_ds.MultiModalX(img1, img2, gt, **hp)
    np.random.seed(11)
    ours = vitcnn_b200.MultiModalX(img1, img2, gt, device=device, **hp)
    assert np.array_equal(np.asarray(theirs.indices), np.asarray(ours.indices))
    return theirs, ours


def test_reference_val_loop_and_loader_mirror(ref_mods):
    """val() of the reference over its own MultiModalX + DataLoader, and vitcnn_b200.val() over the device-resident
    dataset's loader(): same batches in the same order (shuffle included), same accuracy; eval mode and training
    mode (the mode train() validates in)."""
    import vitcnn_b200
    _, ref_ds, ref_mu = ref_mods
    C1, C2, P, K = 32, 1, 7, 6
    img1, img2, gt = R.synthetic_scene(36, 44, C1, C2, K, seed=8)
    theirs, ours = _datasets(ref_ds, img1, img2, gt, P, DEV)
    net = _net(C1, C2, P, K, dropout=0.0)
    # shuffled loaders under the same seed deliver the same samples
    torch.manual_seed(3)
    a = [(d, d2, t) for d, d2, t in torch.utils.data.DataLoader(theirs, batch_size=50, shuffle=True)]
    torch.manual_seed(3)
    b = [(d, d2, t) for d, d2, t in ours.loader(50, shuffle=True)]
    assert len(a) == len(b) == len(ours.loader(50))
    for (d, d2, t), (e, e2, u) in zip(a, b):
        assert torch.equal(d, e.cpu()) and torch.equal(d2, e2.cpu()) and torch.equal(t, u.cpu())
    for mode in ("eval", "train"):
        getattr(net, mode)()
        want = ref_mu.val(net, torch.utils.data.DataLoader(theirs, batch_size=64), device=torch.device(DEV), supervision="full")
        getattr(net, mode)()
        mine = vitcnn_b200.val(net, ours.loader(64), device=torch.device(DEV), supervision="full")
        stock = vitcnn_b200.val(net, torch.utils.data.DataLoader(ours, batch_size=64), device=torch.device(DEV))
        assert 0.0 <= want <= 1.0
        if mode == "eval":
            assert mine == want and stock == want
        else:      # batch statistics: every call also moves the running averages, values stay deterministic
            assert abs(mine - want) <= 1e-12 and abs(stock - want) <= 1e-12
    with pytest.raises(ValueError):
        vitcnn_b200.val(net, ours.loader(64), device=torch.device(DEV), supervision="semi")


class _Recorder(ref_import.NullDisplay):
    def __init__(self):
        self.lines = []

    def line(self, *a, **k):
        self.lines.append({key: np.array(v, dtype=np.float64) for key, v in k.items() if key in ("X", "Y")})
        return "win"


def _train_run(train_fn, ref_ds, sd, img1, img2, gt, train_gt, val_gt, C1, C2, P, K, epochs, with_val, tmp):
    import vitcnn_b200
    from vitcnn_b200.utils import seed_torch
    os.makedirs(tmp, exist_ok=True)
    cwd = os.getcwd()
    os.chdir(tmp)
    try:
        seed_torch(5)
        hp = dict(n_classes=K, n_bands=(C1, C2), ignored_labels=[0], dataset="synthetic", device=torch.device(DEV),
                  patch_size=P, epoch=epochs, batch_size=32)
        net, opt, crit, hp = vitcnn_b200.get_model("ViT-CNN", **hp)
        net.load_state_dict(sd)
        tr = ref_ds.MultiModalX(img1, img2, train_gt, **hp)
        va = ref_ds.MultiModalX(img1, img2, val_gt, **hp)
        tl = torch.utils.data.DataLoader(tr, batch_size=hp["batch_size"], shuffle=True)
        vl = torch.utils.data.DataLoader(va, batch_size=hp["batch_size"])
        rec = _Recorder()
        best = train_fn("x", 0, (C1, C2), net, opt, crit, tl, hp["epoch"], scheduler=hp["scheduler"], display_iter=2,
                        device=hp["device"], display=rec, val_loader=vl if with_val else None, supervision=hp["supervision"])
        files = sorted(os.path.relpath(os.path.join(d, f), tmp)[:-len(f)] + f[19:] for d, _, fs in os.walk(tmp) for f in fs)
        lr = opt.param_groups[0]["lr"]
    finally:
        os.chdir(cwd)
    return best, rec.lines, files, lr, copy.deepcopy(net.state_dict())


@pytest.mark.parametrize("with_val", [True, False])
def test_reference_train_loop_and_mirror_agree(ref_mods, tmp_path, with_val):
    """Three epochs of the reference's unmodified train() -- called exactly as main.py:478 does, with its own
    MultiModalX / DataLoader(shuffle=True), CrossEntropyLoss(weight), optim.Adam, StepLR, validation in training
    mode, checkpoints -- over our module; vitcnn_b200.train() (same signature) must pick the same best epoch,
    write the same checkpoint names, plot the same curves and return the same weights, up to the kernels' own
    run-to-run reproducibility (fp32 atomics in the LayerNorm / pos-embed gradient sums)."""
    import vitcnn_b200
    _, ref_ds, ref_mu = ref_mods
    C1, C2, P, K, epochs = 32, 1, 7, 6, 3
    img1, img2, gt = R.synthetic_scene(40, 56, C1, C2, K, seed=3)
    rng = np.random.default_rng(0)
    pick = rng.random(gt.shape) < 0.12
    train_gt, val_gt = np.where(pick, gt, 0), np.where(~pick & (rng.random(gt.shape) < 0.05), gt, 0)
    torch.manual_seed(0)
    sd = ViTCNNRef(C1, C2, patch_size=P, num_classes=K).state_dict()
    args = (ref_ds, sd, img1, img2, gt, train_gt, val_gt, C1, C2, P, K, epochs, with_val)
    best_r, lines_r, files_r, lr_r, last_r = _train_run(ref_mu.train, *args, str(tmp_path / "ref"))
    best_o, lines_o, files_o, lr_o, last_o = _train_run(vitcnn_b200.train, *args, str(tmp_path / "ours"))
    best_2, lines_2, _, _, last_2 = _train_run(vitcnn_b200.train, *args, str(tmp_path / "ours2"))
    assert files_r == files_o and len(files_r) >= 2 and lr_r == lr_o
    assert len(lines_r) == len(lines_o) > 0
    for a, b in zip(lines_r, lines_o):
        assert a.keys() == b.keys()
        for k in a:
            assert a[k].shape == b[k].shape and np.allclose(a[k], b[k], rtol=2e-3, atol=2e-4), k
    assert best_r.keys() == best_o.keys() == last_r.keys()

    # Yardstick: the backward kernels sum the LayerNorm / pos-embed gradients with fp32 atomics, so two runs of the
    # SAME loop differ in the last bits, and Adam's normalised update amplifies that over the steps.  The mirror must
    # be as close to the reference's loop as it is to its own repeat (the two loops execute the same launches).
    def dist(a, b):
        num = sum(float((a[k].float() - b[k].float()).pow(2).sum()) for k in a if a[k].is_floating_point())
        den = sum(float(a[k].float().pow(2).sum()) for k in a if a[k].is_floating_point())
        return (num / den) ** 0.5

    self_best, self_last = dist(best_o, best_2), dist(last_o, last_2)
    assert dist(best_r, best_o) <= max(4.0 * self_best, 2e-4), (dist(best_r, best_o), self_best)
    assert dist(last_r, last_o) <= max(4.0 * self_last, 2e-4), (dist(last_r, last_o), self_last)
    assert dist(best_r, best_o) <= 2e-2                      # ... and small in absolute terms
    for k in best_r:                                         # integer state (num_batches_tracked) is exact
        if not best_r[k].is_floating_point():
            assert torch.equal(best_r[k], best_o[k]) and torch.equal(last_r[k], last_o[k]), k
    # the best epoch is the same one: without validation the rule keeps the epoch with the HIGHEST mean loss
    # (model_utils.py:1015), i.e. the first, so the returned weights are far from the final ones in both loops
    if not with_val:
        assert dist(best_r, last_r) > 20 * dist(best_r, best_o) and dist(best_o, last_o) > 20 * dist(best_r, best_o)


def test_train_mirror_signature_and_errors():
    """main.py:478 calls train(savename, run, bands, net, optimizer, criterion, loader, epoch, ...) positionally."""
    import inspect
    import vitcnn_b200
    names = list(inspect.signature(vitcnn_b200.train).parameters)
    assert names == ["savename", "run", "bands", "net", "optimizer", "criterion", "data_loader", "epoch", "scheduler",
                     "display_iter", "device", "display", "val_loader", "supervision"]
    with pytest.raises(Exception, match="Missing criterion"):
        vitcnn_b200.train("x", 0, None, None, None, None, [], 1)

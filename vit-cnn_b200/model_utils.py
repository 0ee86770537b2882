"""Host-side mirror of the reference's model factory and evaluation loops for the ViT-CNN
path: same names, argument meaning and error behaviour as ``model_utils.py`` in the
reference, so its callers (main.py:422-500) run unchanged.

* get_model  - model_utils.py:47-68 (head), :206-218 (the FICNN_VIT template the missing
               'ViT-CNN' branch follows), :488-511 (unknown-name KeyError, common tail)
* test       - model_utils.py:1067-1132, computed by ViTCNN.predict_scene on the device
* val        - model_utils.py:1135-1158, vectorised
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
import torch.optim as optim

from .model import SCENE_CHUNK, ViTCNN

MODEL_NAMES = ("ViT-CNN",)


def get_model(name, **kwargs):
    """Instantiate a model with the reference's conventions.

    Returns ``(model, optimizer, criterion, kwargs)``; ``kwargs`` gains the defaults the
    reference's loops read later (patch_size, center_pixel, batch_size, epoch, applyPCA,
    scheduler, supervision, augmentation flags)."""
    device = kwargs.setdefault("device", torch.device("cpu"))
    n_classes = kwargs["n_classes"]
    (n_bands, n_bands2) = kwargs["n_bands"]
    weights = torch.ones(n_classes)
    weights[torch.LongTensor(kwargs["ignored_labels"])] = 0.0
    weights = weights.to(device)
    weights = kwargs.setdefault("weights", weights)
    kwargs["dataset"]  # required key, as in the reference (model_utils.py:67)

    if name == "ViT-CNN":
        kwargs.setdefault("patch_size", 11)
        center_pixel = True
        kwargs.setdefault("applyPCA", False)  # full bands (north star); PCA path is out of scope
        if kwargs["applyPCA"] == True:  # noqa: E712  (mirrors the reference's comparison)
            raise ValueError("ViT-CNN runs on the full band set: applyPCA must be False")
        model = ViTCNN(n_bands, n_bands2, 32, patch_size=kwargs["patch_size"], patch_size_vit=1,
                       num_patches=kwargs["patch_size"] * kwargs["patch_size"], nheads=4, num_layers=2,
                       num_classes=n_classes, dropout=0.01)
        lr = kwargs.setdefault("lr", 0.001)
        optimizer = optim.Adam(model.parameters(), lr=lr)
        criterion = nn.CrossEntropyLoss(weight=kwargs["weights"])
        kwargs.setdefault("epoch", 128)
        kwargs.setdefault("batch_size", 64)
    else:
        raise KeyError("{} model is unknown.".format(name))

    model = model.to(device)
    kwargs.setdefault("scheduler", torch.optim.lr_scheduler.StepLR(optimizer, step_size=30, gamma=0.9))
    kwargs.setdefault("supervision", "full")
    kwargs.setdefault("flip_augmentation", False)
    kwargs.setdefault("radiation_augmentation", False)
    kwargs.setdefault("mixture_augmentation", False)
    kwargs["center_pixel"] = center_pixel
    return model, optimizer, criterion, kwargs


def test(run, net, img1, img2, hyperparams):
    """Full-scene inference with the reference's signature and return value: float64
    ``probs[H, W, n_classes]`` holding the raw logits of the window centred on each pixel and
    exact zeros elsewhere (model_utils.py:1084,1127-1129).  ``run`` is only a progress label
    in the reference."""
    net.eval()
    patch_size = hyperparams["patch_size"]
    center_pixel = hyperparams["center_pixel"]
    device = hyperparams["device"]
    n_classes = hyperparams["n_classes"]
    if hyperparams["applyPCA"] == True:  # noqa: E712
        raise ValueError("ViT-CNN runs on the full band set: applyPCA must be False")
    if not center_pixel:
        raise ValueError("ViT-CNN predicts the centre pixel of each window (center_pixel=True)")
    if not isinstance(net, ViTCNN) or net.patch_size != patch_size or net.num_classes != n_classes:
        raise ValueError("hyperparams do not describe this network")
    from .scene import predict_scene_host
    t1 = torch.as_tensor(np.ascontiguousarray(img1, dtype=np.float32))
    t2 = torch.as_tensor(np.ascontiguousarray(img2, dtype=np.float32))
    chunk = int(hyperparams.get("scene_chunk", SCENE_CHUNK))
    logits_map, _ = predict_scene_host(net, t1, t2, stride=hyperparams["test_stride"], chunk=chunk, device=device)
    return logits_map.numpy().astype(np.float64)


def val(net, data_loader, device="cpu", supervision="full"):
    """Top-1 accuracy where predictions that fall in ``ignored_labels`` are skipped
    (model_utils.py:1135-1158); like the reference it does not switch the network's mode."""
    if supervision != "full":
        raise ValueError('supervision mode "{}" is unknown.'.format(supervision))
    ignored = torch.as_tensor(sorted(data_loader.dataset.ignored_labels), dtype=torch.long, device=device)
    accuracy = torch.zeros((), dtype=torch.long, device=device)
    total = torch.zeros((), dtype=torch.long, device=device)
    for data, data2, target in data_loader:
        with torch.no_grad():
            data, data2, target = data.to(device), data2.to(device), target.to(device)
            output = net(data, data2)
            if isinstance(output, tuple):
                output = output[0]
            pred = torch.argmax(output, dim=1).view(-1)
            keep = ~torch.isin(pred, ignored)
            accuracy += ((pred == target.view(-1)) & keep).sum()
            total += keep.sum()
    return float(accuracy.item()) / float(total.item())


def save_model(savename, model, model_name, dataset_name, train_state, type, **kwargs):
    """Checkpoint writer with the reference's path and file-name scheme (model_utils.py:1047-1064):
    ``./checkpoints/<model_name>/<dataset_name>/<train_state>/<type>/<time><savename>_run{run}_epoch{epoch}_{metric:.2f}.pth``
    holding ``model.state_dict()`` (the keys of ViTCNN are those of the oracle / reference-style module, so
    ``main.py --restore`` -> ``model.load_state_dict(torch.load(path))`` works in both directions).  Anything that
    is not an ``nn.Module`` is pickled with joblib like upstream.  Returns the path written."""
    import datetime
    import os
    model_dir = "./checkpoints/" + model_name + "/" + dataset_name + "/" + train_state + "/" + type + "/"
    time_str = datetime.datetime.now().strftime("%Y_%m_%d_%H_%M_%S")
    os.makedirs(model_dir, exist_ok=True)
    if isinstance(model, torch.nn.Module):
        filename = time_str + savename + "_run{run}_epoch{epoch}_{metric:.2f}".format(**kwargs)
        path = model_dir + filename + ".pth"
        torch.save(model.state_dict(), path)
    else:
        import joblib
        path = model_dir + time_str + ".pkl"
        joblib.dump(model, path)
    return path

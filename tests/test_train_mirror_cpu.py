"""Host logic of the loop mirrors that needs no GPU: signature of train(), the DataLoader-exact batch order of
MultiModalX.loader(), the checkpoint folder name, the running-mean-loss window of the reference's plot."""
import inspect

import numpy as np
import pytest
import torch

from oracle import ref_import


def test_train_signature_is_the_reference_one():
    import vitcnn_b200
    ours = list(inspect.signature(vitcnn_b200.train).parameters.items())
    assert [n for n, _ in ours][:8] == ["savename", "run", "bands", "net", "optimizer", "criterion", "data_loader", "epoch"]
    if ref_import.available():
        _, _, ref_mu = ref_import.import_reference()
        theirs = list(inspect.signature(ref_mu.train).parameters.items())
        assert [n for n, _ in ours] == [n for n, _ in theirs]
        for (_, a), (_, b) in zip(ours, theirs):
            assert (a.default is inspect.Parameter.empty) == (b.default is inspect.Parameter.empty)
            if a.default is not inspect.Parameter.empty:
                assert a.default == b.default
        for fn in ("val", "test", "get_model", "save_model"):
            assert list(inspect.signature(getattr(vitcnn_b200.model_utils, fn)).parameters) == \
                list(inspect.signature(getattr(ref_mu, fn)).parameters), fn


@pytest.mark.parametrize("n,bs", [(23, 7), (64, 64), (1, 5), (130, 32)])
def test_loader_order_equals_the_stock_dataloader(n, bs):
    from vitcnn_b200.datasets import batch_indices
    ds = torch.utils.data.TensorDataset(torch.arange(n))
    for shuffle in (False, True):
        for seed in (0, 9):
            torch.manual_seed(seed)
            want = [b[0].tolist() for b in torch.utils.data.DataLoader(ds, batch_size=bs, shuffle=shuffle)]
            after_w = torch.rand(3)
            torch.manual_seed(seed)
            got = [b.tolist() for b in batch_indices(n, bs, shuffle)]
            after_g = torch.rand(3)
            assert got == want and torch.equal(after_w, after_g)      # same draws from the global generator


def test_camel_to_snake_matches_reference():
    from vitcnn_b200.utils import camel_to_snake
    names = ["ViTCNN", "MFT", "FusAtNet", "Early_fusion_CNN", "S2ENet", "moco_based_NNCNet", "ResNet18", "a", "ABc"]
    if ref_import.available():
        ref_utils, _, _ = ref_import.import_reference(with_model_utils=False)
        for n in names:
            assert camel_to_snake(n) == ref_utils.camel_to_snake(n), n
    assert camel_to_snake("ViTCNN") == "vi_tcnn"


def test_validate_xy():
    from vitcnn_b200.ops import validate_xy
    ok = torch.tensor([[5, 5], [14, 24]], dtype=torch.int32)
    validate_xy(ok, 20, 30, 11, True)
    validate_xy(torch.zeros(0, 2, dtype=torch.int32), 20, 30, 11, True)
    for bad in ([[4, 5]], [[5, 25]], [[15, 5]], [[-1, 5]]):
        with pytest.raises(ValueError):
            validate_xy(torch.tensor(bad, dtype=torch.int32), 20, 30, 11, True)
    validate_xy(torch.tensor([[9, 19]], dtype=torch.int32), 20, 30, 11, False)
    with pytest.raises(ValueError):
        validate_xy(torch.tensor([[10, 19]], dtype=torch.int32), 20, 30, 11, False)

"""-m gpu: the tcgen05 token-stage kernel (csrc/tokens_tc.cu) against the mma.sync kernel
(csrc/transformer.cu) on the same fused feature buffer and the same parameter blob, and against the
fp32 oracle through the module (tests/test_gpu_model.py covers that path for every P <= 11).
Both kernels compute in bf16 with fp32 accumulation, so they agree far inside the 2e-2 oracle
tolerance; the scatter (out_index) and the first-maximum argmax must be identical in behaviour."""
import pytest
import torch

from tests.test_gpu_model import make_pair

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
KERNEL_TOL = 4e-3    # two bf16 pipelines with different summation orders


def _features(n, P, seed):
    from vitcnn_b200 import ops
    g = torch.Generator().manual_seed(seed)
    f = torch.rand(8, ops.sps_rows(n, P), 8, generator=g) * 1.5      # post-ReLU stem outputs: non-negative
    return f.to(torch.bfloat16).to(DEV)


@pytest.mark.parametrize("cfg", [(11, 16, 300), (11, 16, 1), (11, 16, 2), (9, 16, 77), (7, 12, 130), (5, 4, 515), (8, 5, 40),
                                 (10, 8, 33), (1, 3, 9), (2, 64, 12)])
def test_tc_kernel_matches_mma_sync_kernel(cfg):
    from vitcnn_b200 import ops
    P, K, n = cfg
    _, net = make_pair(16, 1, P, K, seed=P)
    blob = net.pack_for_inference()["tparams"]
    f = _features(n, P, seed=n)
    want = ops.tokens_forward(f, blob, n, P, K)
    got = ops.tokens_forward_tc(f, blob, n, P, K)
    torch.cuda.synchronize()
    assert torch.isfinite(got).all()
    err = (got - want).abs().max().item() / want.abs().max().item()
    assert err <= KERNEL_TOL, err


def test_tc_kernel_exact_softmax_path_for_large_attention_weights():
    """The kernel drops the softmax row maximum when a static bound on |q.k| (weight norms) rules out
    overflow of 2^s; large q/k weights must select the two-pass path and still agree."""
    from vitcnn_b200 import ops
    P, K, n = 11, 16, 150
    _, net = make_pair(16, 1, P, K, seed=9)
    with torch.no_grad():
        net.blocks[0].attn.qkv.weight[:64] *= 60.0      # q and k rows: scores of several hundred
        net.blocks[1].attn.qkv.weight[:64] *= 60.0      # ... and for the cls query of the last block
    blob = net.pack_for_inference()["tparams"]
    f = _features(n, P, seed=7)
    want = ops.tokens_forward(f, blob, n, P, K)
    got = ops.tokens_forward_tc(f, blob, n, P, K)
    torch.cuda.synchronize()
    assert torch.isfinite(got).all() and torch.isfinite(want).all()
    err = (got - want).abs().max().item() / want.abs().max().item()
    assert err <= 2e-2, err     # near one-hot attention: a bf16 rounding of a score moves a weight noticeably


def test_tc_kernel_scatter_and_argmax():
    from vitcnn_b200 import ops
    P, K, n = 11, 16, 200
    _, net = make_pair(16, 1, P, K, seed=3)
    blob = net.pack_for_inference()["tparams"]
    f = _features(n, P, seed=1)
    perm = torch.randperm(n + 50, generator=torch.Generator().manual_seed(0))[:n].to(DEV)
    logits = torch.zeros(n + 50, K, device=DEV)
    amap = torch.full((n + 50,), 255, dtype=torch.uint8, device=DEV)
    ops.tokens_forward_tc(f, blob, n, P, K, out_index=perm, logits=logits, argmax_map=amap)
    dense = ops.tokens_forward_tc(f, blob, n, P, K)
    torch.cuda.synchronize()
    assert torch.equal(logits[perm], dense)                       # same kernel, same numbers, scattered rows
    untouched = torch.ones(n + 50, dtype=torch.bool, device=DEV)
    untouched[perm] = False
    assert (logits[untouched] == 0).all() and (amap[untouched] == 255).all()
    assert torch.equal(amap[perm].long(), dense.argmax(1))        # first maximum, like np.argmax


def test_tc_kernel_is_deterministic_and_rejects_large_patches():
    from vitcnn_b200 import ops
    P, K, n = 11, 16, 700
    _, net = make_pair(16, 1, P, K, seed=4)
    blob = net.pack_for_inference()["tparams"]
    f = _features(n, P, seed=2)
    a = ops.tokens_forward_tc(f, blob, n, P, K).clone()
    b = ops.tokens_forward_tc(f, blob, n, P, K)
    torch.cuda.synchronize()
    assert torch.equal(a, b)
    with pytest.raises(RuntimeError):
        ops.tokens_forward_tc(_features(4, 13, 0), blob, 4, 13, K)

"""Time the two token-stage kernels on the same input (CUDA events, L2-cold inputs larger than L2)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vitcnn_b200
from vitcnn_b200 import ops

P, K, n = int(os.environ.get("P", 11)), 16, int(os.environ.get("N", 32768))
dev = "cuda:0"
torch.manual_seed(0)
net = vitcnn_b200.ViTCNN(144, 1, patch_size=P, num_classes=K).to(dev).eval()
blob = net.pack_for_inference()["tparams"]
f = (torch.rand(8, ops.sps_rows(n, P), 8, device=dev) * 1.5).to(torch.bfloat16)
scratch = torch.empty(ops._lib.lib().vc_tokens_tc_scratch_bytes(n), dtype=torch.uint8, device=dev)
variants = (("mma.sync", lambda: ops.tokens_forward(f, blob, n, P, K)),
            ("tcgen05", lambda: ops.tokens_forward_tc(f, blob, n, P, K, scratch=scratch)))
if os.environ.get("ONLY"):
    variants = tuple(v for v in variants if v[0] == os.environ["ONLY"])
for name, fn in variants:
    for _ in range(3):
        out = fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"{name:9s} [{os.environ.get('VITCNN_TC_KERNEL', 'default')}] P={P} n={n}: {ms:.3f} ms  ({9.8e6 * n / ms / 1e9:.1f} TFLOP/s algorithmic at P=11 counts)  finite={torch.isfinite(out).all().item()}")

"""Time the weight-gradient kernel on the training shapes (n = 4096, P = 11) -- tuning aid."""
import os, sys, torch
sys.path.insert(0, ".")
import vitcnn_b200
from vitcnn_b200 import ops, _lib
DEV = "cuda:0"
n, P = int(os.environ.get("N", "4096")), 11
RT = ops.sps_rows(n, P)
RTt = (n * 128 + 127) // 128 * 128
cases = [  # name, SA, SB, taps, shift_on_a, rows_mode, flops/sample
    ("conv_h1", 16, 18, 9, False, False, 2 * 121 * 128 * 144 * 9),
    ("conv_h2", 16, 8, 9, True, False, 2 * 121 * 64 * 128 * 9),
    ("conv_h3", 8, 4, 9, True, False, 2 * 121 * 32 * 64 * 9),
    ("lidar_3", 4, 2, 9, False, False, 2 * 121 * 32 * 16 * 9),
    ("lidar_2", 2, 2, 9, True, False, 2 * 121 * 16 * 8 * 9),
    ("fusion", 8, 4, 1, True, False, 2 * 121 * 32 * 64),
    ("tok_qkv", 12, 6, 1, False, True, 2 * 122 * 96 * 32),
    ("tok_fc2", 4, 18, 1, False, True, 2 * 122 * 32 * 128),
]
for name, SA, SB, taps, soa, rows_mode, fl in cases:
    rows = RTt if rows_mode else RT
    A = torch.randn(SA, rows, 8, device=DEV).to(torch.bfloat16)
    B = torch.randn(SB, rows, 8, device=DEV).to(torch.bfloat16)
    out = torch.zeros(128 * 256 * 9, device=DEV)
    ws = torch.empty(_lib.lib().vc_wgrad_workspace_bytes(SB, taps), dtype=torch.uint8, device=DEV)
    def run():
        if rows_mode:
            ops.wgrad_sps(A, B, rows // 128, 0, taps, soa, out, SA * 8, SB * 8, SB * 8, 1, 0, workspace=ws)
        else:
            ops.wgrad_sps(A, B, n, P, taps, soa, out, SA * 8, SB * 8, SB * 8 * taps, taps, 1, workspace=ws)
    for _ in range(3):
        run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        run()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 100
    gb = (SA + SB) * rows * 16 / 1e9
    print(f"{name:8s} rows={os.environ.get('VC_WGRAD_ROWS','auto'):>4s} {us:8.1f} us  {fl * n / us / 1e6:7.1f} TFLOP/s(alg)  {gb / us * 1e6:7.0f} GB/s")

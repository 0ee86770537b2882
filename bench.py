"""Headline benchmark: full-scene ViT-CNN inference on a synthetic Houston2013-shaped raster
(349 x 1905, 144 + 1 bands, 16 classes, patch 11, stride 1), BASELINE.json configs[1].

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = N_gpus scenes, every scene split into N_gpus row bands (one per rank, halo rows
included, no collective), i.e. one scene's worth of windows per GPU per step (weak scaling).
Prints ONE JSON line (rank 0).  `value` = pixels/s with rasters resident in HBM; `e2e` = the
same through the public host-buffer API (pinned H2D of the band + D2H of its logits/argmax
every step); `roofline` = the dominant kernel (HSI stem conv 1, tcgen05) from a CUDA-event
profiled pass of the same step; `cpu_baseline` = the fp32 oracle on the host cores on a
bounded sample.  --impl reference times that CPU path alone (the reference has no GPU code
of its own and ships no ViT-CNN source: the oracle port is its stand-in).
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, W, C1, C2, K, P = 349, 1905, 144, 1, 16, 11
WORKLOAD = "houston2013_full_scene_inference_p11_stride1"


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        d = json.load(open(path))
        return d["hbm_gbs"], d["bf16_tflops"], d.get("bf16_tflops_sustained", d["bf16_tflops"]), "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_oracle_rate(seconds: float, threads: int):
    """fp32 oracle model fed by the oracle's restatement of test()'s batch assembly, on the host
    cores, for about `seconds` of work on windows taken from the centre rows of the scene."""
    from oracle import data_ref as R
    from oracle.model_ref import ViTCNNRef
    torch.set_num_threads(threads)
    rng = np.random.default_rng(0)
    rows = P + 3                                   # a band of 4 window rows
    img1 = rng.random((rows, W, C1), dtype=np.float32)
    img2 = rng.random((rows, W, C2), dtype=np.float32)
    torch.manual_seed(0)
    ref = ViTCNNRef(C1, C2, patch_size=P, num_classes=K).eval()
    corners = R.sliding_window_corners((rows, W), 1, (P, P))
    probs = np.zeros((rows, W, K))
    done, t0 = 0, time.perf_counter()
    with torch.no_grad():
        for s in range(0, len(corners), 64):       # batch 64 = the reference's default
            chunk = corners[s:s + 64]
            h, l = R.gather_corners(img1, img2, chunk, P)
            out = ref(torch.from_numpy(h), torch.from_numpy(l)).numpy()
            for (x, y), o in zip(chunk, out):
                probs[x + P // 2, y + P // 2] += o
            done += len(chunk)
            if time.perf_counter() - t0 > seconds:
                break
    dt = time.perf_counter() - t0
    n_windows = (H - P + 1) * (W - P + 1)
    return done / dt * (H * W) / n_windows, done, dt   # pixels/s of a full scene at this window rate


def train_flops_per_sample():
    """Algorithmic FLOPs of one training sample: forward + data-gradient +
    weight-gradient GEMMs = 3 x forward, minus the data gradient of the two input convs."""
    from oracle.model_ref import forward_flops
    f = forward_flops(C1, C2, P, K)
    first = 2 * P * P * 9 * (C1 * 128 + C2 * 8)
    return 3 * f["total"] - first


def bench_train(args, rank, world, dev, dist, barrier):
    """Training throughput, BASELINE.json configs[2]: batch data-parallel, global batch 4096
    (4096 / N per GPU), patches gathered on the device from the resident raster, weighted CE,
    backward, all-reduce of one flat fp32 bucket (NCCL), Adam.  samples/s = global batch / step."""
    import vitcnn_b200
    from vitcnn_b200 import _lib
    from vitcnn_b200.train import Trainer
    rng = np.random.default_rng(1)
    img1 = torch.from_numpy(rng.random((H, W, C1), dtype=np.float32)).to(dev)
    img2 = torch.from_numpy(rng.random((H, W, C2), dtype=np.float32)).to(dev)
    gt_h = rng.integers(1, K, size=(H, W)).astype(np.int64)
    gt = torch.from_numpy(gt_h).to(dev)
    torch.manual_seed(0)
    net = vitcnn_b200.ViTCNN(C1, C2, patch_size=P, num_classes=K, dropout=0.01).to(dev)   # reference default (model_utils.py:213)
    w = torch.ones(K)
    w[0] = 0
    tr = Trainer(net, lr=1e-3, weights=w, use_graph=not args.no_graph)
    per = args.train_batch // world
    p = P // 2
    nbatch = 8
    xy_h = [torch.from_numpy(np.stack([rng.integers(p + 1, H - p - 1, per), rng.integers(p + 1, W - p - 1, per)], 1)
                             .astype(np.int32)).pin_memory() for _ in range(nbatch)]
    xy_d = [t.to(dev) for t in xy_h]
    it = [0]

    def step():
        tr.step(img1, img2, gt, xy_d[it[0] % nbatch])
        it[0] += 1

    loss_h = torch.empty(2).pin_memory()

    def e2e_step():
        xy = xy_h[it[0] % nbatch].to(dev, non_blocking=True)
        loss = tr.step(img1, img2, gt, xy)
        loss_h.copy_(loss, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        it[0] += 1

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    steps = max(args.steps, 5) * 4
    for _ in range(max(args.warmup, 3) + 4):      # includes the eager steps before the graph capture
        step()
    l0 = _lib.lib().vc_launch_count()
    ms = timed(step, steps)
    launches = _lib.lib().vc_launch_count() - l0
    if tr.use_graph and tr.launches_per_step:      # replays do not pass through the library's counter
        launches = tr.launches_per_step * steps
    for _ in range(2):
        e2e_step()
    ms_e2e = timed(e2e_step, steps)
    tr.use_graph = False                          # per-kernel-class events need eager launches
    step()
    _lib.profile_begin()
    step()
    prof = _lib.profile_end()
    hbm, tf_burst, tf_sust, which = peaks()
    fl = train_flops_per_sample() * per
    total_ms = sum(v[0] for v in prof.values())
    return {"metric": "train_samples_per_s", "value": args.train_batch * steps / (ms / 1e3), "unit": "samples/s",
            "ms_per_step": ms / steps, "steps": steps, "global_batch": args.train_batch, "per_gpu_batch": per,
            "parallelism": f"dp{world}", "scaling": "strong", "dtype": "bf16", "optimizer": "Adam(lr=1e-3)",
            "loss": "CrossEntropy(weight)", "dropout": 0.01, "gpu_launches": int(launches), "cuda_graph": not args.no_graph,
            "e2e": {"value": args.train_batch * steps / (ms_e2e / 1e3), "unit": "samples/s", "ms_per_step": ms_e2e / steps,
                    "h2d_bytes_per_step": int(per * 8), "d2h_bytes_per_step": 8},
            "roofline": {"bound": "tensor", "achieved": fl / (ms / steps / 1e3) / 1e12, "peak": tf_sust, "unit": "TFLOP/s",
                         "frac": fl / (ms / steps / 1e3) / 1e12 / tf_sust, "peak_source": which + " bf16 sustained",
                         "scope": "whole training step, algorithmic FLOPs (3 x forward - input-conv dgrad)",
                         "breakdown_ms": {k: round(v[0], 3) for k, v in prof.items() if v[1]},
                         "profiled_step_ms": total_ms}}


def cpu_train_rate(seconds: float, threads: int, batch: int = 64):
    """Training throughput of the fp32 oracle on the host cores: the reference's own loop body
    (model_utils.py:908-945: forward, weighted CE, backward, Adam) at its default batch size 64 on
    patches cut from a synthetic Houston-shaped raster, for about `seconds` of work."""
    from oracle import data_ref as R
    from oracle.model_ref import ViTCNNRef
    import torch.nn.functional as F
    torch.set_num_threads(threads)
    rng = np.random.default_rng(2)
    rows = 3 * P
    img1 = rng.random((rows, W, C1), dtype=np.float32)
    img2 = rng.random((rows, W, C2), dtype=np.float32)
    gt = rng.integers(1, K, size=(rows, W)).astype(np.int64)
    torch.manual_seed(0)
    ref = ViTCNNRef(C1, C2, patch_size=P, num_classes=K).train()
    opt = torch.optim.Adam(ref.parameters(), lr=1e-3)
    w = torch.ones(K)
    w[0] = 0
    p = P // 2
    done, t0 = 0, time.perf_counter()
    while True:
        xy = np.stack([rng.integers(p + 1, rows - p - 1, batch), rng.integers(p + 1, W - p - 1, batch)], 1)
        h, l, y = R.gather_centers(img1, img2, gt, xy, P)
        loss = F.cross_entropy(ref(torch.from_numpy(h), torch.from_numpy(l)), torch.from_numpy(y), w)
        opt.zero_grad()
        loss.backward()
        opt.step()
        done += batch
        if time.perf_counter() - t0 > seconds:
            break
    dt = time.perf_counter() - t0
    return done / dt, done, dt


def run_reference(args, rank):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    for _ in range(max(args.warmup, 0)):
        cpu_oracle_rate(1.0, threads)
    vals, samples = [], 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        v, n, _ = cpu_oracle_rate(8.0, threads)
        vals.append(v)
        samples += n
    dt = time.perf_counter() - t0
    v = float(np.mean(vals))
    tv, tn, tdt = cpu_train_rate(8.0, threads)
    line = {"impl": "reference", "metric": "full_scene_inference_pixels_per_s", "value": v, "unit": "pixels/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "note": "fp32 oracle port on host cores (reference ships no ViT-CNN source)"},
            "cpu_baseline": {"value": v, "unit": "pixels/s", "cores": threads, "kind": "port",
                             "sample": f"{samples} windows of a {P + 3}-row band, batch 64, extrapolated to the scene"},
            "e2e": {"value": v, "unit": "pixels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "train": {"metric": "train_samples_per_s", "value": tv, "unit": "samples/s", "cores": threads, "kind": "port",
                      "sample": f"{tn} samples ({tdt:.1f} s), batch 64 (the reference's default), fp32 oracle + torch Adam"}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--chunk", type=int, default=int(os.environ.get("VITCNN_CHUNK", "131072")))
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--windows", type=int, default=0, help="profiling aid: only the first N windows of the band")
    ap.add_argument("--no-cpu", action="store_true", help="profiling aid: skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true", help="profiling aid: skip the e2e leg")
    ap.add_argument("--no-train", action="store_true", help="skip the training-throughput leg")
    ap.add_argument("--no-graph", action="store_true", help="training leg: eager launches instead of one CUDA graph per step")
    ap.add_argument("--no-infer", action="store_true", help="profiling aid: training leg only")
    ap.add_argument("--train-batch", type=int, default=4096, help="GLOBAL batch of the training leg (configs[2])")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    args.warmup = max(args.warmup, 3)

    import torch.distributed as dist
    import vitcnn_b200
    from vitcnn_b200 import _lib
    from vitcnn_b200.utils import row_band_ranges, window_starts

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    rng = np.random.default_rng(0)
    img1_h = torch.from_numpy(rng.random((H, W, C1), dtype=np.float32)).pin_memory()
    img2_h = torch.from_numpy(rng.random((H, W, C2), dtype=np.float32)).pin_memory()
    torch.manual_seed(0)
    net = vitcnn_b200.ViTCNN(C1, C2, patch_size=P, num_classes=K).to(dev).eval()
    img1, img2 = img1_h.to(dev), img2_h.to(dev)
    nx, ny = len(window_starts(H, P, 1)), len(window_starts(W, P, 1))
    first, count = row_band_ranges(nx, ny, world)[rank]
    if args.windows:
        count = min(count, args.windows)
    scenes = world                                        # scenes per step (weak scaling)
    logits_map = torch.zeros(H, W, K, dtype=torch.float32, device=dev)
    argmax_map = torch.zeros(H, W, dtype=torch.uint8, device=dev)

    def step():
        for _ in range(scenes):
            net.predict_scene(img1, img2, chunk=args.chunk, window_range=(first, count), logits_map=logits_map,
                              argmax_map=argmax_map)

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    train_line = None
    if not args.no_train:
        train_line = bench_train(args, rank, world, dev, dist, barrier)
        torch.cuda.empty_cache()
    if args.no_infer:
        if rank == 0:
            print(json.dumps({"train": train_line}), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return
    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = _lib.lib().vc_launch_count()
    ms = timed(step, args.steps)
    launches = _lib.lib().vc_launch_count() - l0
    clocks = sampler.stop()
    pixels_per_step = scenes * H * W * (count * world / (nx * ny) if args.windows else 1.0)
    value = pixels_per_step * args.steps / (ms / 1e3)

    # ---- end to end through the host-buffer API --------------------------------------------
    x_first, x_last = first // ny, (first + count - 1) // ny          # window rows of this band
    band = slice(x_first, x_last + P)                                 # raster rows incl. halo
    out_rows = slice(x_first + P // 2, x_last + P // 2 + 1)           # rows this band writes
    lg_h = torch.empty(H, W, K, dtype=torch.float32).pin_memory()
    am_h = torch.empty(H, W, dtype=torch.uint8).pin_memory()

    def e2e_step():
        for _ in range(scenes):
            vitcnn_b200.predict_scene_host(net, img1_h, img2_h, rank=rank, world=world, chunk=args.chunk,
                                           logits_out=lg_h, argmax_out=am_h)
    if args.no_e2e or args.windows:
        ms_e2e = float("nan")
    else:
        for _ in range(2):
            e2e_step()
        ms_e2e = timed(e2e_step, args.steps)
    band_rows = band.stop - band.start
    h2d = scenes * band_rows * W * (C1 + C2) * 4
    d2h = scenes * (out_rows.stop - out_rows.start) * W * (K * 4 + 1)
    e2e_val = pixels_per_step * args.steps / (ms_e2e / 1e3)

    # ---- roofline of the dominant kernel: CUDA-event profiled pass of the same step ----------
    _lib.profile_begin()
    step()
    prof = _lib.profile_end()
    total_ms = sum(v[0] for v in prof.values())
    c1_ms, c1_n = prof["conv_h1"]
    hbm, tf_burst, tf_sust, which = peaks()
    nwin = count * scenes                                             # windows of this rank in one step
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r01_conv1_traffic.json")
    if os.path.isfile(tpath):       # dram bytes per launch from the committed ncu capture (same chunk size only)
        tj = json.load(open(tpath))
        if tj.get("chunk_windows") == min(args.chunk, count):
            traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
    T_ = P * P + 1
    token_flops = (2 * P * P * 64 * 32 + 2 * (2 * T_ * 32 * 96 + 2 * T_ * 32 * 32 + 2 * 4 * 2 * T_ * T_ * 8 + 2 * 2 * T_ * 32 * 128)
                   + 2 * 32 * K)                                      # fusion + 2 blocks + head
    S1 = (C1 + 15) // 16 * 2
    # Shared stem (vc_scene_infer): the first `depth` HSI convs run once per scene on 31 x 31 (depth 1: 15 x 15)
    # blocks as (2L+1)^2 border-class variants and the windows gather their stem output from them.  The work of
    # those kernels is the work they execute on the blocks (taps that leave a window are not issued), NOT the
    # per-window figure of SURVEY App. D -- the sharing is an algorithmic saving, not tensor throughput.
    pk = net.pack_for_inference()
    band_rows_dev = (first + count - 1) // ny - first // ny + P      # raster rows predict_scene hands to the library
    n_chunks = -(-count // args.chunk)
    chunk_eff = -(-count // n_chunks)
    ws_bytes = _lib.lib().vc_scene_workspace_bytes(ctypes.byref(pk["struct"]), band_rows_dev, W, chunk_eff)
    depth = _lib.lib().vc_scene_shared_depth(ctypes.byref(pk["struct"]), band_rows_dev, W, chunk_eff, count, ws_bytes)
    Bb = 31 if depth >= 2 else 15
    step_b = Bb - 2 * max(depth, 1)
    nb = ((band_rows_dev - Bb + step_b - 1) // step_b + 1) * ((W - Bb + step_b - 1) // step_b + 1) if depth else 0
    taps1, taps2 = 49, 169                                            # issued (variant, tap) pairs: (2+3+2)^2, (2+3+3+3+2)^2
    blk_px = Bb * Bb
    w_c1 = 2.0 * 128 * C1 * taps1 * blk_px * nb * scenes if depth >= 1 else 2.0 * P * P * 128 * C1 * 9 * nwin
    w_c2 = 2.0 * 64 * 128 * taps2 * blk_px * nb * scenes if depth >= 2 else 2.0 * P * P * 64 * 128 * 9 * nwin
    g_slices = {0: S1, 1: 16, 2: 8, 3: 4}[depth]
    kernels = {   # class -> (name, bound, work of this rank in one step, unit scale, peak)
        "conv_h1": (("conv_sps_tc2_kernel x 9 border-class variants on %d x %d scene blocks (HSI conv 1, tcgen05)" % (Bb, Bb))
                    if depth >= 1 else "conv_sps_tc2_kernel (HSI stem conv1, tcgen05)", "tensor", w_c1, 1e12, tf_sust, "TFLOP/s"),
        "conv_h2": ("conv_sps_tc_kernel x 25 variants, multi-plane input (HSI conv 2, tcgen05)" if depth >= 2
                    else "conv_sps_tc_kernel (HSI stem conv2, tcgen05)", "tensor", w_c2, 1e12, tf_sust, "TFLOP/s"),
        "tokens": ("tokens_tc_kernel + tokens_tail_kernel (token stage, tcgen05)" if (82 <= P * P + 1 <= 128 and os.environ.get("VITCNN_TOKENS_IMPL") != "0")
                   else "transformer_fwd_kernel (token stage, mma.sync)", "tensor", float(token_flops) * nwin, 1e12, tf_sust, "TFLOP/s"),
        # HBM bytes that must move: the bf16 SPS rows written per window (gathered stem slices + LiDAR slices) and
        # the same bytes read (variant planes / raster; re-reads across overlapping windows are L2 hits)
        "pack": ("border_gather_kernel (stem variants -> window SPS) + pack_strip_kernel (LiDAR)" if depth
                 else "pack_strip_kernel (TMA-staged patch gather -> bf16 SPS)", "hbm",
                 float((g_slices + 2) * (P + 1) * (P + 1) * 16) * nwin
                 + (float((2 * depth + 1) ** 2 * g_slices * 16 * nb * (Bb + 1) ** 2) * scenes if depth
                    else float((C1 + C2) * 4 * H * W) * scenes * count / (nx * ny)), 1e9, hbm, "GB/s"),
    }

    def entry(cls):
        name, bound, work, scale, peak, unit = kernels[cls]
        t_ms, n_l = prof[cls]
        ach = work / (t_ms / 1e3) / scale if t_ms > 0 else 0.0
        return {"kernel": name, "bound": bound, "achieved": ach, "peak": peak, "unit": unit, "frac": ach / peak,
                "peak_source": which + (" bf16 sustained" if bound == "tensor" else " HBM copy"),
                "launches": n_l, "avg_launch_ms": t_ms / max(n_l, 1), "share_of_step": t_ms / total_ms if total_ms else None}

    dominant = max(("conv_h1", "tokens"), key=lambda c: prof[c][0])   # the single kernel with the largest share
    roofline = entry(dominant)
    roofline["traffic"] = traffic if (dominant == "conv_h1" and depth == 0) else None
    ttpath = os.path.join(ROOT, "profiles", "r01_tokens_traffic.json")
    if dominant == "tokens" and "tokens_tc" in roofline["kernel"] and os.path.isfile(ttpath):
        tj = json.load(open(ttpath))      # dram bytes per launch from the committed ncu capture (same chunk size only)
        if tj.get("chunk_windows") == min(args.chunk, count):
            roofline["traffic"] = tj["dram_bytes_read"] + tj["dram_bytes_write"]
    roofline["shared_stem_depth"] = int(depth)
    if dominant == "tokens":   # what actually bounds this kernel: transcendentals, not the tensor pipe
        sfu_ms = (4 * T_ * 128 + T_ * 128) * nwin / (16.0 * 148 * clocks.get("sm_mhz", 1965.0) * 1e6) * 1e3 if clocks.get("sm_mhz") else None
        roofline["limiter"] = {"unit": "SFU (MUFU ex2 / tanh), 16 results per clock and SM (tools/probe/mufu_probe.cu)",
                               "floor_ms_per_step": sfu_ms, "frac_of_floor": (sfu_ms / prof["tokens"][0]) if sfu_ms else None}
    roofline["breakdown_ms"] = {k: round(v[0], 3) for k, v in prof.items() if v[1]}
    roofline["other_kernels"] = [dict(entry(c), traffic=(traffic if (c == "conv_h1" and depth == 0) else None)) for c in kernels if c != dominant]

    if rank == 0:
        cpu_v, cpu_n, cpu_dt = (float("nan"), 0, 0.0) if args.no_cpu else cpu_oracle_rate(args.cpu_seconds,
                                                                                         os.cpu_count() or 1)
        line = {"metric": "full_scene_inference_pixels_per_s", "value": value, "unit": "pixels/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": {"workload": WORKLOAD, "scene": [H, W, C1, C2], "classes": K, "patch": P,
                           "windows_per_scene": nx * ny, "scenes_per_step": scenes, "sharding": "row-band",
                           "chunk_windows": args.chunk, "l2": "inputs larger than L2 (385 MB raster)"},
                "windows_per_s": nx * ny * scenes * args.steps / (ms / 1e3),
                "clocks": clocks, "gpu_launches": int(launches),
                "e2e": {"value": e2e_val, "unit": "pixels/s", "h2d_bytes_per_step": int(h2d),
                        "d2h_bytes_per_step": int(d2h), "ms_per_step": ms_e2e / args.steps},
                "roofline": roofline,
                "cpu_baseline": {"value": cpu_v, "unit": "pixels/s", "cores": os.cpu_count() or 1, "kind": "port",
                                 "sample": f"{cpu_n} windows ({cpu_dt:.1f} s) of a {P + 3}-row band, batch 64, "
                                           "extrapolated to the scene"}}
        if train_line is not None:
            if not args.no_cpu:
                tv, tn, tdt = cpu_train_rate(min(args.cpu_seconds, 8.0), os.cpu_count() or 1)
                train_line["cpu_baseline"] = {"value": tv, "unit": "samples/s", "cores": os.cpu_count() or 1, "kind": "port",
                                              "sample": f"{tn} samples ({tdt:.1f} s), batch 64, fp32 oracle + torch Adam"}
            line["train"] = train_line
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

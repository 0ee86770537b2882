"""Headline benchmark: full-scene ViT-CNN inference on a synthetic Houston2013-shaped raster
(349 x 1905, 144 + 1 bands, 16 classes, patch 11, stride 1), BASELINE.json configs[1].

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                  [--workload houston|muufl|augsburg] [--patch 7|9|11|15]        (configs[3], configs[4])
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = N_gpus scenes, every scene split into N_gpus row bands (one per rank, halo rows
included, no collective), i.e. one scene's worth of windows per GPU per step (weak scaling).
Prints ONE JSON line (rank 0):
  value        pixels/s with the rasters resident in HBM (CUDA events, max over ranks);
  e2e          the same through the public host-buffer API (predict_scene_host: pinned H2D of every band,
               D2H of its logits / argmax map, every step); also the host time spent queueing a step, the cudaMalloc
               count of the timed region and the upload-only time of the same bytes (what the link alone allows);
  strong       ONE scene split over the N ranks: ms per scene (max over ranks, barrier to barrier), device and
               end to end, and an order-independent integer checksum of the assembled maps against the map one
               GPU computes alone (row bands must reproduce it bit for bit);
  roofline     the dominant kernel (the token stage) from a CUDA-event profiled pass of the same step,
               other_kernels = every other kernel class of the step;
  cpu_baseline the reference's own test() loop (from baseline/_ref when present) driving the fp32 oracle model
               on the host cores, on a bounded sample;
  train        BASELINE.json's second metric (configs[2]): data-parallel training, global batch 4096.
--impl reference times that CPU path alone: the reference has no GPU code of its own and ships no ViT-CNN
model source, so its loops (unmodified) drive the oracle port of the model.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import tempfile
import threading
import time
from types import SimpleNamespace

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# synthetic shapes of BASELINE.json: H, W, HSI bands, LiDAR bands, classes
WORKLOADS = {
    "houston": dict(H=349, W=1905, C1=144, C2=1, K=16, name="houston2013"),
    "muufl": dict(H=325, W=220, C1=64, C2=2, K=12, name="muufl"),
    "augsburg": dict(H=332, W=485, C1=180, C2=1, K=8, name="augsburg"),
}


def make_cfg(workload: str, patch: int) -> SimpleNamespace:
    c = SimpleNamespace(**WORKLOADS[workload])
    c.P = int(patch)
    c.workload = f"{c.name}_full_scene_inference_p{c.P}_stride1"
    return c


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


# ---------------------------------------------------------------------------------------------------------
# CPU arm: the reference's loops over the oracle model (the only places bench.py touches oracle/)
# ---------------------------------------------------------------------------------------------------------
def _reference_modules():
    """(utils, datasets, model_utils) of the unmodified reference (baseline/_ref or /root/reference), or None."""
    from oracle import ref_import
    if not ref_import.available():
        return None
    try:
        return ref_import.import_reference()
    except Exception:
        return None


def cpu_scene_rate(c, seconds: float, threads: int):
    """Scene inference on the host cores: the reference's own ``model_utils.test()`` (model_utils.py:1067-1132,
    sliding_window -> grouper -> batch assembly -> net -> scatter) driving the fp32 oracle model, on a band of
    rows from a synthetic raster of the workload's shape, repeated until about `seconds` have passed.  Falls back
    to the oracle's restatement of that loop when the reference files are not there.  Returns (pixels/s of a full
    scene at the measured window rate, windows done, seconds, kind)."""
    from oracle import data_ref as R
    from oracle.model_ref import ViTCNNRef
    torch.set_num_threads(threads)
    rng = np.random.default_rng(0)
    rows = c.P + 1                                   # a band of 2 window rows
    img1 = rng.random((rows, c.W, c.C1), dtype=np.float32)
    img2 = rng.random((rows, c.W, c.C2), dtype=np.float32)
    torch.manual_seed(0)
    ref = ViTCNNRef(c.C1, c.C2, patch_size=c.P, num_classes=c.K).eval()
    mods = _reference_modules()
    n_band = (rows - c.P + 1) * (c.W - c.P + 1)
    done, t0 = 0, time.perf_counter()
    if mods is not None:
        kind = "reference-loop+port-model"
        hp = dict(patch_size=c.P, center_pixel=True, batch_size=64, device=torch.device("cpu"), n_classes=c.K,
                  applyPCA=False, test_stride=1)
        with open(os.devnull, "w") as sink:           # test() drives a tqdm bar
            err, sys.stderr = sys.stderr, sink
            try:
                while True:
                    mods[2].test(0, ref, img1, img2, hp)
                    done += n_band
                    if time.perf_counter() - t0 > seconds:
                        break
            finally:
                sys.stderr = err
    else:
        kind = "port"
        corners = R.sliding_window_corners((rows, c.W), 1, (c.P, c.P))
        probs = np.zeros((rows, c.W, c.K))
        with torch.no_grad():
            while time.perf_counter() - t0 <= seconds:
                for s in range(0, len(corners), 64):       # batch 64 = the reference's default
                    chunk = corners[s:s + 64]
                    h, l = R.gather_corners(img1, img2, chunk, c.P)
                    out = ref(torch.from_numpy(h), torch.from_numpy(l)).numpy()
                    for (x, y), o in zip(chunk, out):
                        probs[x + c.P // 2, y + c.P // 2] += o
                    done += len(chunk)
    dt = time.perf_counter() - t0
    n_windows = (c.H - c.P + 1) * (c.W - c.P + 1)
    return done / dt * (c.H * c.W) / n_windows, done, dt, kind


def cpu_train_rate(c, seconds: float, threads: int, batch: int = 64):
    """BASELINE.json configs[0]: ONE epoch of the reference's own ``train()`` (model_utils.py:854-1045) over its own
    ``MultiModalX`` + ``DataLoader(shuffle=True)`` with the fp32 oracle model, CrossEntropyLoss(weight), Adam(1e-3),
    batch 64, on the host cores; the number of labelled pixels is sized for about `seconds` of work.  Falls back to
    the loop body on the oracle's gather when the reference files are not there.  Returns (samples/s, samples,
    seconds, kind)."""
    from oracle import data_ref as R
    from oracle.model_ref import ViTCNNRef
    import torch.nn.functional as F
    torch.set_num_threads(threads)
    rng = np.random.default_rng(2)
    rows = 4 * c.P
    img1 = rng.random((rows, c.W, c.C1), dtype=np.float32)
    img2 = rng.random((rows, c.W, c.C2), dtype=np.float32)
    torch.manual_seed(0)
    ref = ViTCNNRef(c.C1, c.C2, patch_size=c.P, num_classes=c.K).train()
    opt = torch.optim.Adam(ref.parameters(), lr=1e-3)
    w = torch.ones(c.K)
    w[0] = 0
    p = c.P // 2
    mods = _reference_modules()
    if mods is not None:
        # size the epoch: time two batches of the loop body first
        probe = np.stack([rng.integers(p + 1, rows - p - 1, batch), rng.integers(p + 1, c.W - p - 1, batch)], 1)
        h, l, y = R.gather_centers(img1, img2, np.ones((rows, c.W), np.int64), probe, c.P)
        for k in range(5):                               # two warm-up batches, three timed ones
            if k == 2:
                t0 = time.perf_counter()
            loss = F.cross_entropy(ref(torch.from_numpy(h), torch.from_numpy(l)), torch.from_numpy(y), w)
            opt.zero_grad()
            loss.backward()
            opt.step()
        per_batch = (time.perf_counter() - t0) / 3
        n_samples = int(max(2, min(200, round(seconds / max(per_batch, 1e-3)))) * batch)
        gt = np.zeros((rows, c.W), np.int64)
        inner = [(x, yy) for x in range(p + 1, rows - p - 1) for yy in range(p + 1, c.W - p - 1)]
        pick = rng.choice(len(inner), size=min(n_samples, len(inner)), replace=False)
        for k in pick:
            gt[inner[k]] = 1 + (k % (c.K - 1))
        hp = dict(dataset="synthetic", patch_size=c.P, ignored_labels=[0], flip_augmentation=False,
                  radiation_augmentation=False, mixture_augmentation=False, center_pixel=True, supervision="full",
                  applyPCA=False)
        ds = mods[1].MultiModalX(img1, img2, gt, **hp)
        loader = torch.utils.data.DataLoader(ds, batch_size=batch, shuffle=True)
        from oracle.ref_import import NullDisplay
        cwd = os.getcwd()
        with tempfile.TemporaryDirectory() as tmp, open(os.devnull, "w") as sink:
            os.chdir(tmp)                               # train() writes ./checkpoints/...
            err, out, sys.stderr, sys.stdout = sys.stderr, sys.stdout, sink, sink
            try:
                t0 = time.perf_counter()
                mods[2].train("bench", 0, (c.C1, c.C2), ref, opt, torch.nn.CrossEntropyLoss(weight=w), loader, 1,
                              scheduler=None, display_iter=10 ** 9, device=torch.device("cpu"), display=NullDisplay(),
                              val_loader=None, supervision="full")
                dt = time.perf_counter() - t0
            finally:
                sys.stderr, sys.stdout = err, out
                os.chdir(cwd)
        return len(ds) / dt, len(ds), dt, "reference-loop+port-model"
    gt = rng.integers(1, c.K, size=(rows, c.W)).astype(np.int64)
    done, t0 = 0, time.perf_counter()
    while True:
        xy = np.stack([rng.integers(p + 1, rows - p - 1, batch), rng.integers(p + 1, c.W - p - 1, batch)], 1)
        h, l, y = R.gather_centers(img1, img2, gt, xy, c.P)
        loss = F.cross_entropy(ref(torch.from_numpy(h), torch.from_numpy(l)), torch.from_numpy(y), w)
        opt.zero_grad()
        loss.backward()
        opt.step()
        done += batch
        if time.perf_counter() - t0 > seconds:
            break
    dt = time.perf_counter() - t0
    return done / dt, done, dt, "port"


def run_reference(args, c, rank):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    for _ in range(max(args.warmup, 0)):
        cpu_scene_rate(c, 1.0, threads)
    vals, samples, kind = [], 0, "port"
    t0 = time.perf_counter()
    for _ in range(args.steps):
        v, n, _, kind = cpu_scene_rate(c, 8.0, threads)
        vals.append(v)
        samples += n
    dt = time.perf_counter() - t0
    v = float(np.mean(vals))
    tv, tn, tdt, tkind = cpu_train_rate(c, 8.0, threads)
    line = {"impl": "reference", "metric": "full_scene_inference_pixels_per_s", "value": v, "unit": "pixels/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": c.workload, "scene": [c.H, c.W, c.C1, c.C2], "classes": c.K, "patch": c.P,
                       "note": "the reference's own test() loop (unmodified, baseline/_ref) over the fp32 oracle port of the "
                               "model on the host cores (the reference ships no ViT-CNN model source)"},
            "cpu_baseline": {"value": v, "unit": "pixels/s", "cores": threads, "kind": kind,
                             "sample": f"{samples} windows of a {c.P + 1}-row band, batch 64, extrapolated to the scene"},
            "e2e": {"value": v, "unit": "pixels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "train": {"metric": "train_samples_per_s", "value": tv, "unit": "samples/s", "cores": threads, "kind": tkind,
                      "sample": f"one epoch of {tn} samples ({tdt:.1f} s) through the reference's train(), batch 64, fp32 oracle + "
                                "torch Adam (BASELINE.json configs[0])"}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------
def train_flops_per_sample(c):
    """Algorithmic FLOPs of one training sample: forward + data-gradient +
    weight-gradient GEMMs = 3 x forward, minus the data gradient of the two input convs."""
    from oracle.model_ref import forward_flops
    f = forward_flops(c.C1, c.C2, c.P, c.K)
    first = 2 * c.P * c.P * 9 * (c.C1 * 128 + c.C2 * 8)
    return 3 * f["total"] - first


def bench_train(args, c, rank, world, dev, dist, barrier):
    """Training throughput, BASELINE.json configs[2]: batch data-parallel, global batch 4096
    (4096 / N per GPU), patches gathered on the device from the resident raster, weighted CE,
    backward, all-reduce of one flat fp32 bucket (NCCL), Adam.  samples/s = global batch / step.
    `dp_check`: after the timed steps every rank's flat parameter bucket must hold the same bits (integer
    checksum, min == max over ranks), and the loss of the step is reported next to the loss one GPU gets on the
    SAME global batch (BatchNorm statistics are per rank, as in torch DDP without SyncBN, so the two differ by the
    batch-statistics noise only)."""
    import vitcnn_b200
    from vitcnn_b200 import _lib
    from vitcnn_b200.train import Trainer
    H, W, C1, C2, K, P = c.H, c.W, c.C1, c.C2, c.K, c.P
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
    xy_all = [np.stack([rng.integers(p + 1, H - p - 1, args.train_batch), rng.integers(p + 1, W - p - 1, args.train_batch)], 1)
              .astype(np.int32) for _ in range(nbatch)]                          # GLOBAL batches, identical on every rank
    xy_h = [torch.from_numpy(np.ascontiguousarray(a[rank * per:(rank + 1) * per])).pin_memory() for a in xy_all]
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
    # ---- data-parallel correctness: identical replicas, loss against one GPU on the same global batch ----
    flat = tr.state.flat
    cs = flat.view(torch.int32).to(torch.int64).sum().reshape(1)
    cs_min, cs_max = cs.clone(), cs.clone()
    if world > 1:
        dist.all_reduce(cs_min, op=dist.ReduceOp.MIN)
        dist.all_reduce(cs_max, op=dist.ReduceOp.MAX)
    k = it[0] % nbatch
    tr.use_graph = False                          # eager launches from here on (loss read-back, per-kernel events)
    loss_dp = tr.step(img1, img2, gt, xy_d[k])[0:1].clone()
    it[0] += 1
    if world > 1:
        dist.all_reduce(loss_dp, op=dist.ReduceOp.SUM)
        loss_dp /= world
    dp_check = {"ranks": world, "steps": int(it[0]), "replicas_identical": bool(cs_min.item() == cs_max.item()),
                "param_checksum": int(cs.item()), "loss_mean_over_ranks": float(loss_dp.item()),
                "batchnorm": "statistics per rank (torch DDP semantics without SyncBN)"}
    if world > 1 and rank == 0:
        # the same step on ONE GPU: a replica with rank 0's pre-step weights is not kept, so compare on the forward
        # loss of the CURRENT weights over the whole global batch (BatchNorm batch statistics over 4096 instead of 4096 / N)
        from vitcnn_b200.train import ce_loss
        net.train()
        with torch.no_grad():
            xy_g = torch.from_numpy(xy_all[k]).to(dev)
            lg, lab = tr.state.forward_gather(img1, img2, gt, xy_g)
            l1, _ = ce_loss(lg, lab, tr.weights, want_grad=False)
            lg2, lab2 = tr.state.forward_gather(img1, img2, gt, xy_g[:per].contiguous())
            l2, _ = ce_loss(lg2, lab2, tr.weights, want_grad=False)
        dp_check["loss_one_gpu_global_batch"] = float(l1[0].item())
        dp_check["loss_rank0_shard_same_weights"] = float(l2[0].item())
    step()
    _lib.profile_begin()
    step()
    prof = _lib.profile_end()
    hbm, tf_burst, tf_sust, which = peaks()
    fl = train_flops_per_sample(c) * per
    total_ms = sum(v[0] for v in prof.values())
    breakdown = {k2: round(v[0], 3) for k2, v in prof.items() if v[1]}
    breakdown["allreduce"] = round(tr.allreduce_ms(), 4)      # the one collective of the step, timed alone (0 for one rank)
    tr.close()                                                # graphs with NCCL kernels must be gone before the group is destroyed
    return {"metric": "train_samples_per_s", "value": args.train_batch * steps / (ms / 1e3), "unit": "samples/s",
            "ms_per_step": ms / steps, "steps": steps, "global_batch": args.train_batch, "per_gpu_batch": per,
            "parallelism": f"dp{world}", "scaling": "strong", "dtype": "bf16", "optimizer": "Adam(lr=1e-3)",
            "loss": "CrossEntropy(weight)", "dropout": 0.01, "gpu_launches": int(launches), "cuda_graph": not args.no_graph,
            "e2e": {"value": args.train_batch * steps / (ms_e2e / 1e3), "unit": "samples/s", "ms_per_step": ms_e2e / steps,
                    "h2d_bytes_per_step": int(per * 8), "d2h_bytes_per_step": 8},
            "dp_check": dp_check,
            "roofline": {"bound": "tensor", "achieved": fl / (ms / steps / 1e3) / 1e12, "peak": tf_sust, "unit": "TFLOP/s",
                         "frac": fl / (ms / steps / 1e3) / 1e12 / tf_sust, "peak_source": which + " bf16 sustained",
                         "scope": "whole training step, algorithmic FLOPs (3 x forward - input-conv dgrad)",
                         "breakdown_ms": breakdown, "profiled_step_ms": total_ms}}


def int_checksum(t: torch.Tensor) -> torch.Tensor:
    """Order-independent checksum of a tensor's BITS (int64 sum of its 32-bit words / bytes)."""
    if t.dtype == torch.float32:
        return t.contiguous().view(torch.int32).to(torch.int64).sum()
    return t.to(torch.int64).sum()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="houston", choices=sorted(WORKLOADS), help="raster shape (BASELINE.json configs[3])")
    ap.add_argument("--patch", type=int, default=11, help="patch size P (configs[4]: 7 / 9 / 11 / 15)")
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
    # a rank that gets stuck (a collective its peers never enter, a wedged teardown) must not hold the node: dump every
    # thread's stack to stderr and leave after VITCNN_BENCH_WATCHDOG seconds (default 15 min; the default run needs ~1.5)
    import faulthandler
    faulthandler.dump_traceback_later(int(os.environ.get("VITCNN_BENCH_WATCHDOG", "900")), exit=True)
    c = make_cfg(args.workload, args.patch)
    H, W, C1, C2, K, P = c.H, c.W, c.C1, c.C2, c.K, c.P
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, c, rank)
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

    def one_scene():
        net.predict_scene(img1, img2, chunk=args.chunk, window_range=(first, count), logits_map=logits_map,
                          argmax_map=argmax_map)

    def step():
        for _ in range(scenes):
            one_scene()

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
        train_line = bench_train(args, c, rank, world, dev, dist, barrier)
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

    # ---- end to end through the host-buffer API: scene after scene, streaming ------------------
    x_first, x_last = first // ny, (first + count - 1) // ny          # window rows of this band
    band = slice(x_first, x_last + P)                                 # raster rows incl. halo
    out_rows = slice(x_first + P // 2, x_last + P // 2 + 1)           # rows this band writes
    outs = [(torch.empty(H, W, K, dtype=torch.float32).pin_memory(), torch.empty(H, W, dtype=torch.uint8).pin_memory())
            for _ in range(2)]                                        # double-buffered host results
    e2e_it = [0]

    def e2e_scene(sync):
        lg_h, am_h = outs[e2e_it[0] & 1]
        e2e_it[0] += 1
        vitcnn_b200.predict_scene_host(net, img1_h, img2_h, rank=rank, world=world, chunk=args.chunk, logits_out=lg_h,
                                       argmax_out=am_h, sync=sync)

    host_s = [0.0, 0]         # host time spent queueing (time.perf_counter around the calls: no synchronisation inside)

    def e2e_step():          # every scene: H2D of the band from pinned memory, kernels, D2H of its maps; the next scene's
        t0 = time.perf_counter()
        for _ in range(scenes):   # upload overlaps this one's tail (the final synchronize of the timed region drains all)
            e2e_scene(False)
        host_s[0] += time.perf_counter() - t0
        host_s[1] += 1
    if args.no_e2e or args.windows:
        ms_e2e = float("nan")
    else:
        for _ in range(max(args.warmup, 3)):     # the caching allocator must have seen the streaming depth before the timed region
            e2e_step()
        host_s[0], host_s[1] = 0.0, 0
        st0 = torch.cuda.memory_stats(dev)
        ms_e2e = timed(e2e_step, args.steps)
        st1 = torch.cuda.memory_stats(dev)
        e2e_allocs = {k: (st1.get(k, 0) - st0.get(k, 0)) / args.steps for k in ("num_device_alloc", "num_device_free", "num_alloc_retries")}
    host_queue_ms = host_s[0] / max(host_s[1], 1) * 1e3
    if args.no_e2e or args.windows:
        e2e_allocs = {}
    band_rows = band.stop - band.start
    # what the host link alone allows: the same pinned -> HBM uploads with no kernels in between (the e2e leg cannot be
    # faster than this; on the round-2 boxes a Houston scene's 386 MB take ~34 ms, i.e. the leg is link-bound once the
    # device needs less than that)
    def h2d_only():
        for _ in range(scenes):
            img1_h[band].to(dev, non_blocking=True)
            img2_h[band].to(dev, non_blocking=True)
    ms_h2d = float("nan") if (args.no_e2e or args.windows) else timed(h2d_only, args.steps)
    h2d = scenes * band_rows * W * (C1 + C2) * 4
    d2h = scenes * (out_rows.stop - out_rows.start) * W * (K * 4 + 1)
    e2e_val = pixels_per_step * args.steps / (ms_e2e / 1e3)

    # ---- strong scaling: ONE scene over the N ranks -------------------------------------------------
    strong = None
    if not args.windows:
        ms_strong = timed(one_scene, args.steps)
        if not args.no_e2e:
            for _ in range(max(args.warmup, 3)):     # the single-scene plan has its own sub-bands and staging size: warm it up too
                e2e_scene(True)
        ms_strong_e2e = float("nan") if args.no_e2e else timed(lambda: e2e_scene(True), args.steps)
        logits_map.zero_()
        argmax_map.zero_()
        one_scene()
        cs = torch.stack([int_checksum(logits_map), int_checksum(argmax_map)])
        if world > 1:
            dist.all_reduce(cs, op=dist.ReduceOp.SUM)      # bands are disjoint, everything else is zero
        want = cs.clone()
        if world > 1:
            if rank == 0:                                  # the whole scene on one GPU
                full_l = torch.zeros_like(logits_map)
                full_a = torch.zeros_like(argmax_map)
                net.predict_scene(img1, img2, chunk=args.chunk, logits_map=full_l, argmax_map=full_a)
                want = torch.stack([int_checksum(full_l), int_checksum(full_a)])
                del full_l, full_a
            dist.broadcast(want, src=0)
        strong = {"scenes": 1, "ranks": world, "ms_per_scene": ms_strong / args.steps,
                  "value": H * W * args.steps / (ms_strong / 1e3), "unit": "pixels/s",
                  "e2e_ms_per_scene": ms_strong_e2e / args.steps, "e2e_value": H * W * args.steps / (ms_strong_e2e / 1e3),
                  "map_checksum": [int(v) for v in cs.tolist()], "one_gpu_checksum": [int(v) for v in want.tolist()],
                  "checksum_match": bool(torch.equal(cs, want)),
                  "checksum": "int64 sum of the 32-bit words of the logits map / of the argmax bytes (order independent)"}

    # ---- roofline: CUDA-event profiled pass of the same step, per kernel class ------------------
    _lib.profile_begin()
    step()
    prof = _lib.profile_end()
    total_ms = sum(v[0] for v in prof.values())
    hbm, tf_burst, tf_sust, which = peaks()
    nwin = count * scenes                                             # windows of this rank in one step
    T_ = P * P + 1
    token_flops = (2 * P * P * 64 * 32 + 2 * (2 * T_ * 32 * 96 + 2 * T_ * 32 * 32 + 2 * 4 * 2 * T_ * T_ * 8 + 2 * 2 * T_ * 32 * 128)
                   + 2 * 32 * K)                                      # fusion + 2 blocks + head
    S1 = (C1 + 15) // 16 * 2
    # Shared stem (vc_scene_infer): the first `depth` HSI convs (and, whenever the geometry allows, all three LiDAR
    # convs) run once per call on 31 x 31 (depth 1: 15 x 15) scene blocks as (2L+1)^2 border-class variants; at depth 3
    # the token kernel reads the variant planes itself.  The work of those kernels is the work they EXECUTE on the
    # blocks (taps that leave a window are not issued), NOT the per-window figure of SURVEY App. D -- the sharing is
    # an algorithmic saving, not tensor throughput.
    pk = net.pack_for_inference()
    band_rows_dev = (first + count - 1) // ny - first // ny + P      # raster rows predict_scene hands to the library
    n_chunks = -(-count // args.chunk)
    chunk_eff = -(-count // n_chunks)
    ws_bytes = _lib.lib().vc_scene_workspace_bytes(ctypes.byref(pk["struct"]), band_rows_dev, W, chunk_eff)
    depth = _lib.lib().vc_scene_shared_depth(ctypes.byref(pk["struct"]), band_rows_dev, W, chunk_eff, count, ws_bytes)

    def n_blocks(Bb, D):
        sb = Bb - 2 * D
        return ((band_rows_dev - Bb + sb - 1) // sb + 1) * ((W - Bb + sb - 1) // sb + 1)

    Bb = int(_lib.lib().vc_scene_block(band_rows_dev, W, max(int(depth), 1)))
    nb = n_blocks(Bb, max(depth, 1)) if depth else 0
    pairs = (49, 169, 361)                                            # issued (variant, tap) pairs of conv 1 / 2 / 3: 7^2, 13^2, 19^2
    blk_px = Bb * Bb
    cin_h, cout_h = (C1, 128, 64), (128, 64, 32)
    w_h = [2.0 * cout_h[l] * cin_h[l] * pairs[l] * blk_px * nb * scenes if depth > l
           else 2.0 * P * P * cout_h[l] * cin_h[l] * 9 * nwin for l in range(3)]
    lidar_shared = bool(depth) and P >= 7 and band_rows_dev >= 31 and W >= 31 and os.environ.get("VITCNN_LIDAR_SHARED", "1") != "0"
    cin_l, cout_l = (C2, 8, 16), (8, 16, 32)
    Bl = int(_lib.lib().vc_scene_block(band_rows_dev, W, 3))
    nb_l = n_blocks(Bl, 3) if lidar_shared else 0
    w_l = sum(2.0 * cout_l[l] * cin_l[l] * pairs[l] * Bl * Bl * nb_l * scenes if lidar_shared
              else 2.0 * P * P * cout_l[l] * cin_l[l] * 9 * nwin for l in range(3))
    tc_tokens = 26 <= T_ <= 128 and os.environ.get("VITCNN_TOKENS_IMPL") != "0"
    direct = depth == 3 and tc_tokens
    g_slices = {0: S1, 1: 16, 2: 8, 3: 4}[depth]
    shared_name = "conv_var_kernel: all %d border-class variants per launch, one work unit per (tile, row class) (%s, tcgen05)"
    kernels = {   # class -> (name, bound, work of this rank in one step, unit scale, peak, unit)
        "conv_h1": (("conv_sps_tc2_kernel x 9 border-class variants on %d x %d scene blocks (HSI conv 1, tcgen05 CTA pairs)" % (Bb, Bb))
                    if depth >= 1 else "conv_sps_tc2_kernel (HSI stem conv 1 per window, tcgen05 CTA pairs)", "tensor", w_h[0], 1e12, tf_sust, "TFLOP/s"),
        "conv_h2": (shared_name % (25, "HSI conv 2") if depth >= 2 else "conv_sps_tc_kernel (HSI stem conv 2 per window, tcgen05)",
                    "tensor", w_h[1], 1e12, tf_sust, "TFLOP/s"),
        "conv_h3": (shared_name % (49, "HSI conv 3") if depth >= 3 else "conv_sps_tc_kernel (HSI stem conv 3 per window, tcgen05)",
                    "tensor", w_h[2], 1e12, tf_sust, "TFLOP/s"),
        "conv_lidar": ("LiDAR stem shared at depth 3: conv_sps_tc_kernel x 9 (conv 1) + conv_var_kernel (conv 2, conv 3) (tcgen05)"
                       if lidar_shared else "conv_sps_tc_kernel x 3 (LiDAR stem per window, tcgen05)", "tensor", w_l, 1e12, tf_sust, "TFLOP/s"),
        "tokens": (((("tokens_tc_kernel" if os.environ.get("VITCNN_TC_KERNEL") == "tc" else "tokens_tm_kernel (probabilities / hidden units through TMEM, %s patches in flight per SM)"
                      % ("3" if os.environ.get("VITCNN_TC_KERNEL") == "tm3" else "4"))
                     + " + tokens_tail_kernel (token stage, tcgen05") + (", stem inputs read from the variant planes)" if direct else ")"))
                   if tc_tokens else "transformer_fwd_kernel (token stage, mma.sync)", "tensor", float(token_flops) * nwin, 1e12, tf_sust, "TFLOP/s"),
        # HBM bytes that must move in the gather / packing launches: block packing (fp32 raster read, bf16 SPS written),
        # plus, when the token kernel does not read the planes itself, the per-window SPS rows written and read
        "pack": ("pack_sps_kernel (scene blocks -> bf16 SPS)" + ("" if direct else " + border_gather_kernel (stem variants -> window SPS)")
                 if depth else "pack_strip_kernel (TMA-staged patch gather -> bf16 SPS)", "hbm",
                 (float((C1 * 4 + S1 * 16) * nb * (Bb + 1) ** 2 + (C2 * 4 + 32) * nb_l * (Bl + 1) ** 2) * scenes if depth else 0.0)
                 + (0.0 if direct else float(2 * (g_slices + (4 if lidar_shared else 2)) * (P + 1) * (P + 1) * 16) * nwin), 1e9, hbm, "GB/s"),
    }

    def entry(cls):
        name, bound, work, scale, peak, unit = kernels[cls]
        t_ms, n_l = prof[cls]
        ach = work / (t_ms / 1e3) / scale if t_ms > 0 else 0.0
        return {"kernel": name, "bound": bound, "achieved": ach, "peak": peak, "unit": unit, "frac": ach / peak,
                "peak_source": which + (" bf16 sustained" if bound == "tensor" else " HBM copy"),
                "launches": n_l, "avg_launch_ms": t_ms / max(n_l, 1), "share_of_step": t_ms / total_ms if total_ms else None}

    present = [k for k in kernels if prof[k][1]]
    dominant = max(present, key=lambda k: prof[k][0])                 # the single kernel class with the largest share
    roofline = entry(dominant)
    roofline["traffic"] = None
    ttpath = os.path.join(ROOT, "profiles", "r02_tokens_traffic.json")
    if dominant == "tokens" and tc_tokens and os.path.isfile(ttpath):
        tj = json.load(open(ttpath))      # dram bytes per launch from the committed ncu capture (same launch size and input path only)
        if tj.get("chunk_windows") == min(args.chunk, count) and tj.get("direct") == direct and tj.get("workload") == c.workload:
            roofline["traffic"] = tj["dram_bytes_read"] + tj["dram_bytes_write"]
    roofline["shared_stem_depth"] = int(depth)
    if dominant == "tokens":   # what actually bounds this kernel: transcendentals, not the tensor pipe
        sfu_ms = (4 * T_ * 128 + T_ * 128) * nwin / (16.0 * 148 * clocks.get("sm_mhz", 1965.0) * 1e6) * 1e3 if clocks.get("sm_mhz") else None
        roofline["limiter"] = {"unit": "SFU (MUFU ex2 / tanh), 16 results per clock and SM (tools/probe/mufu_probe.cu)",
                               "floor_ms_per_step": sfu_ms, "frac_of_floor": (sfu_ms / prof["tokens"][0]) if sfu_ms else None}
    roofline["breakdown_ms"] = {k: round(v[0], 3) for k, v in prof.items() if v[1]}
    roofline["other_kernels"] = [dict(entry(k), traffic=None) for k in present if k != dominant]

    if rank == 0:
        ncpu = os.cpu_count() or 1
        cpu_v, cpu_n, cpu_dt, cpu_kind = (float("nan"), 0, 0.0, "port") if args.no_cpu else cpu_scene_rate(c, args.cpu_seconds, ncpu)
        line = {"metric": "full_scene_inference_pixels_per_s", "value": value, "unit": "pixels/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": {"workload": c.workload, "scene": [H, W, C1, C2], "classes": K, "patch": P,
                           "windows_per_scene": nx * ny, "scenes_per_step": scenes, "sharding": "row-band",
                           "chunk_windows": args.chunk,
                           "l2": "inputs larger than L2 (%d MB raster, %d MB of stem variant planes per scene)"
                                 % (H * W * (C1 + C2) * 4 // 10 ** 6, (ws_bytes - _lib.lib().vc_workspace_bytes(chunk_eff, P, C1, C2)) // 10 ** 6)},
                "windows_per_s": nx * ny * scenes * args.steps / (ms / 1e3),
                "clocks": clocks, "gpu_launches": int(launches),
                "e2e": {"value": e2e_val, "unit": "pixels/s", "h2d_bytes_per_step": int(h2d),
                        "d2h_bytes_per_step": int(d2h), "ms_per_step": ms_e2e / args.steps,
                        "host_queue_ms_per_step": host_queue_ms, "cuda_mallocs_per_step": e2e_allocs, "h2d_only_ms_per_step": ms_h2d / args.steps, "h2d_only_gb_per_s": h2d / (ms_h2d / args.steps) / 1e6,
                        "api": "vitcnn_b200.predict_scene_host(sync=False): scene after scene, double-buffered pinned results"},
                "strong": strong,
                "roofline": roofline,
                "cpu_baseline": {"value": cpu_v, "unit": "pixels/s", "cores": ncpu, "kind": cpu_kind,
                                 "sample": f"{cpu_n} windows ({cpu_dt:.1f} s) of a {P + 1}-row band through the reference's test(), "
                                           "batch 64, extrapolated to the scene"}}
        if train_line is not None:
            if not args.no_cpu:
                tv, tn, tdt, tkind = cpu_train_rate(c, min(args.cpu_seconds, 8.0), ncpu)
                train_line["cpu_baseline"] = {"value": tv, "unit": "samples/s", "cores": ncpu, "kind": tkind,
                                              "sample": f"one epoch of {tn} samples ({tdt:.1f} s) through the reference's train(), "
                                                        "batch 64, fp32 oracle + torch Adam (BASELINE.json configs[0])"}
            line["train"] = train_line
        print(json.dumps(line), flush=True)
    if world > 1:
        import gc
        gc.collect()
        torch.cuda.synchronize()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

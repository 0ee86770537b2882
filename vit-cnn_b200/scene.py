"""Host-buffer entry point of full-scene inference: upload the rank's row band (plus halo
rows) from pinned host memory, run the scene kernels, download the band's logits / argmax.
This is the `e2e` path of bench.py and what test() uses; row bands need no collective
(SURVEY.md section 8(e))."""
from __future__ import annotations

import ctypes
import os

import torch

from . import _lib
from .model import SCENE_CHUNK
from .utils import band_geometry, plan_subbands, subband_bounds


@torch.no_grad()
def predict_scene_host(net, img1: torch.Tensor, img2: torch.Tensor, stride: int = 1, rank: int = 0, world: int = 1,
                       chunk: int = SCENE_CHUNK, logits_out: torch.Tensor = None, argmax_out: torch.Tensor = None,
                       device=None, pipeline: int = None, sync: bool = True):
    """img1 f32 [H,W,C1], img2 f32 [H,W,C2] CPU tensors (pinned for async copies).  Writes the
    rows owned by ``rank`` into ``logits_out`` f32 [H,W,K] / ``argmax_out`` uint8 [H,W] (CPU,
    allocated zero-filled when None) and returns them.  Rows no window is centred on are not
    touched.

    ``sync=False`` (streaming: scene after scene, e.g. the N scenes of a bench step) returns as soon as the
    work is queued: the uploads of the NEXT call then overlap the tail of this one instead of waiting for its
    last download.  The outputs are complete once the device is synchronised (``torch.cuda.synchronize``) or
    the returned tensors' ``ready`` event (``predict_scene_host.last_event(device)``) has passed; at most two
    calls are in flight per device (the third waits for the first), pinned outputs must not be reused before
    their call has completed."""
    H, W, _ = img1.shape
    P, K = net.patch_size, net.num_classes
    dev = torch.device(device) if device is not None else net.cls_token.device
    if logits_out is None:
        logits_out = torch.zeros(H, W, K, dtype=torch.float32)
    if argmax_out is None:
        argmax_out = torch.zeros(H, W, dtype=torch.uint8)
    geo = band_geometry(H, W, P, stride, rank, world)
    if geo["count"] == 0:
        return logits_out, argmax_out
    x0, x1 = geo["x0"], geo["x1"]
    if pipeline is None:
        # sub-bands per call.  One scene at a time: 6 (upload / compute / download of consecutive sub-bands overlap inside the
        # call).  Streaming: the next call's first upload already overlaps this call's tail, so fewer, larger sub-bands win (each
        # one re-runs the shared stem's ~25 launches).  Measured, Houston scene on one B200, ms per scene at 1 / 2 / 3 / 4 / 6
        # sub-bands: streaming 41.2 / 40.9 / 42.1 / 42.5 / 42.9, single 48.3 / 45.1 / 45.0 / 44.7 / 44.4 (device only: 40.2)
        # A short band (one of many ranks' share of a scene: 43 window rows at 8 GPUs) goes as ONE sub-band when streaming:
        # every sub-band re-runs the shared stem's launches on whole scene blocks, and with 8 ranks on 16 host cores the
        # queueing itself (41 ms per 8-band step, device 38 ms) was what bounded the end-to-end rate.
        pipeline = int(os.environ.get("VITCNN_PIPELINE", "6" if sync else ("2" if len(geo["xs"]) > 96 else "1")))
    # Software pipeline over `pipeline` sub-bands of the band: the pinned-host -> HBM upload of sub-band
    # k+1 and the HBM -> host download of sub-band k-1 run on copy streams while sub-band k computes,
    # so the PCIe traffic (386 MB up, 42 MB down for a Houston scene) hides behind the kernels.
    xs_rel, P2 = geo["xs"], P // 2
    nrows = len(xs_rel)
    # sub-bands of at least 21 window rows (31 raster rows at P = 11): the shared stem of vc_scene_infer needs
    # 31-row rasters, and shorter sub-bands would spend their time on launch tails.  Dense scenes (stride 1)
    # are cut at whole block rows of that stem (utils.subband_bounds).  Measured on B200, Houston scene
    # (tools/gpu_subbands.sh): 8 equal parts (16 block rows) 60.0 ms; whole block rows at 8 / 7 / 6 / 5 sub-bands
    # 59.4 / 59.4 / 58.2 / 58.2 ms (15 / 15 / 14 / 14 block rows; one uncut band needs 13); a 21-row lead
    # sub-band (shorter first upload) gains nothing (58.4 / 58.5 ms at 5 / 6), so the default is 6 even ones
    nsub = max(1, min(int(pipeline), nrows // 21))
    bounds = [(nrows * k) // nsub for k in range(nsub + 1)]
    spans = list(zip(bounds[:-1], bounds[1:]))         # window rows [a, b) of every sub-band
    if stride == 1 and nsub > 1 and os.environ.get("VITCNN_SUBBAND_SPLIT", "blocks") != "equal":
        # how many sub-bands, and where to cut: utils.plan_subbands (block rows of the stem vs the exposed first upload)
        L = _lib.lib()
        lead_small = os.environ.get("VITCNN_SUBBAND_LEAD", "even") == "small"
        key = (nrows, W, P, len(geo["ys"]), nsub, bool(sync), int(chunk), lead_small, str(dev), net.n_bands, net.n_bands2, net.num_classes)
        best = _PLAN.get(key)
        if best is None:
            best = plan_subbands(nrows, P, nsub, sync,
                                 lambda rows: _shared_depth(net, rows + P - 1, W, rows * len(geo["ys"]), chunk, dev),
                                 lambda h, depth: int(L.vc_scene_block(h, W, depth)), lead_small)
        if best is not None:
            _PLAN[key] = best
            spans = best[1]
            nsub = len(spans)
    with torch.cuda.device(dev):
        main = torch.cuda.current_stream()
        up, down = _copy_streams(dev)
        inflight = _INFLIGHT.setdefault(dev, [])
        if sync:
            up.wait_stream(main)
        elif len(inflight) >= 2:              # bounded look-ahead: staging memory of at most two calls
            up.wait_event(inflight.pop(0))
        if sync:
            for ev in inflight:               # streaming calls still in flight own the staging slots
                up.wait_event(ev)
        # Device staging of this call (rasters of every sub-band, their result maps) comes from a persistent arena with two
        # slots, one per call in flight: a steady stream of calls performs no cudaMalloc (fresh `.to(dev)` / torch.zeros
        # tensors were only reusable after their recorded streams had drained, so the caching allocator kept growing:
        # ~3 cudaMalloc per scene, 2 - 15 ms of host time per call depending on the box)
        subs = []
        for k in range(nsub):
            a, b = spans[k]
            if a == b:
                subs.append(None)
                continue
            xa, xb = x0 + int(xs_rel[a]), x0 + int(xs_rel[b - 1]) + P           # raster rows of the sub-band
            subs.append((xa, xb, xs_rel[a:b] + x0 - xa))
        C1, C2 = img1.shape[2], img2.shape[2]
        views = _stage_views(dev, [(xb - xa, W, C1, C2, K) for (xa, xb, _) in (s for s in subs if s is not None)])
        staged, vi = [], 0
        for k in range(nsub):
            if subs[k] is None:
                staged.append(None)
                continue
            xa, xb, xs_k = subs[k]
            b1, b2, lg, am = views[vi]
            vi += 1
            with torch.cuda.stream(up):
                b1.copy_(img1[xa:xb], non_blocking=True)
                b2.copy_(img2[xa:xb], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(up)
            staged.append((b1, b2, lg, am, ev, xa, xs_k))
        for k in range(nsub):
            if staged[k] is None:
                continue
            b1, b2, lg, am, ev, xa, xs_k = staged[k]
            main.wait_event(ev)
            lg.zero_()
            am.zero_()
            net.predict_scene(b1, b2, stride=stride, chunk=chunk, xs=xs_k, logits_map=lg, argmax_map=am)
            o0, o1 = xa + int(xs_k[0]) + P2, xa + int(xs_k[-1]) + P2 + 1       # map rows with window centres
            l0 = int(xs_k[0]) + P2
            cev = torch.cuda.Event()
            cev.record(main)
            down.wait_event(cev)
            with torch.cuda.stream(down):
                logits_out[o0:o1].copy_(lg[l0:l0 + (o1 - o0)], non_blocking=True)
                argmax_out[o0:o1].copy_(am[l0:l0 + (o1 - o0)], non_blocking=True)
        if sync:
            main.wait_stream(down)
            main.synchronize()
            inflight.clear()
        else:                                 # the downloads are ordered on `down` after the kernels they read from
            ev = torch.cuda.Event()
            ev.record(down)
            inflight.append(ev)
    return logits_out, argmax_out


def last_event(device):
    """Event recorded after the most recent ``sync=False`` call on ``device`` (None when nothing is in flight)."""
    q = _INFLIGHT.get(torch.device(device), [])
    return q[-1] if q else None


predict_scene_host.last_event = last_event
_INFLIGHT = {}


def _shared_depth(net, rows: int, W: int, count: int, chunk: int, dev) -> int:
    """Sharing depth vc_scene_infer picks for a sub-band of ``rows`` raster rows holding ``count`` windows
    (0: per-window stem), asked of the library with the chunking predict_scene will use."""
    if count <= 0:
        return 0
    L = _lib.lib()
    with torch.cuda.device(dev):
        pk = net.pack_for_inference()
    n_chunks = -(-count // int(chunk))
    chunk_eff = -(-count // n_chunks)
    ws = L.vc_scene_workspace_bytes(ctypes.byref(pk["struct"]), rows, W, chunk_eff)
    return max(0, int(L.vc_scene_shared_depth(ctypes.byref(pk["struct"]), rows, W, chunk_eff, count, ws)))


_STREAMS = {}
_PLAN = {}       # sub-band plans by geometry
_ARENA = {}      # device -> [arena uint8 tensor, next slot]


def _stage_views(dev, shapes):
    """Views (raster 1, raster 2, logits map, argmax map) per sub-band into the next slot of the device's staging arena.
    The arena holds two slots (at most two calls are in flight per device); it grows -- after a device synchronise, so
    nothing in flight loses its memory -- when a call needs more than a slot holds."""
    def rnd(n):
        return (n + 255) & ~255
    need = sum(rnd(r * w * c1 * 4) + rnd(r * w * c2 * 4) + rnd(r * w * k * 4) + rnd(r * w) for (r, w, c1, c2, k) in shapes)
    ent = _ARENA.get(dev)
    if ent is None or ent[0].numel() < 2 * need:
        torch.cuda.synchronize(dev)
        ent = _ARENA[dev] = [torch.empty(2 * need, dtype=torch.uint8, device=dev), 0]
    arena, slot = ent
    half = arena.numel() // 2
    ent[1] = slot ^ 1
    off = slot * half
    out = []
    for (r, w, c1, c2, k) in shapes:
        def take(nbytes, dtype, shape):
            nonlocal off
            v = arena[off:off + nbytes].view(dtype).view(shape)
            off += rnd(nbytes)
            return v
        out.append((take(r * w * c1 * 4, torch.float32, (r, w, c1)), take(r * w * c2 * 4, torch.float32, (r, w, c2)),
                    take(r * w * k * 4, torch.float32, (r, w, k)), take(r * w, torch.uint8, (r, w))))
    return out


def _copy_streams(dev):
    st = _STREAMS.get(dev)
    if st is None:
        st = _STREAMS[dev] = (torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev))
    return st

"""Thin torch-tensor wrappers over the C-ABI entry points (device pointers + current stream).
These are the operator-level API: tests, the dataset mirror and the profiling scripts call
the kernels through here; nothing in this file computes."""
from __future__ import annotations

import torch

from . import _lib


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def sps_rows(n: int, P: int) -> int:
    return int(_lib.lib().vc_sps_rows(n, P))


def validate_xy(xy, H: int, W: int, P: int, center_mode: bool = True) -> None:
    """Raise ``ValueError`` unless every P x P window of ``xy`` (int [n,2], centres or top-left corners) lies
    inside the H x W raster -- what the reference's strict-border rule (datasets.py:497-504) and
    ``sliding_window`` (utils.py:374-397) guarantee for their own coordinates.  One device->host read."""
    if xy.numel() == 0:
        return
    if P > H or P > W:
        raise ValueError(f"patch {P} does not fit a {H} x {W} raster")
    lo, hi = xy.amin(0).tolist(), xy.amax(0).tolist()
    off = P // 2 if center_mode else 0
    if lo[0] - off < 0 or lo[1] - off < 0 or hi[0] - off + P > H or hi[1] - off + P > W:
        raise ValueError(f"patch coordinates leave the {H} x {W} raster for P = {P} "
                         f"(rows {lo[0]}..{hi[0]}, columns {lo[1]}..{hi[1]}, center_mode={bool(center_mode)})")


def gather_patches(img1, img2, xy, P, center_mode=True, gt=None, ops=None, validate=True):
    """Exact fp32 patch extraction.  img1 [H,W,C1] f32, img2 [H,W,C2] f32 (CUDA, contiguous),
    xy int32 [n,2] centres (center_mode) or top-left corners.  Returns (hsi [n,C1,P,P],
    lidar [n,C2,P,P], labels int64 [n] or None).  ``ops`` (uint8 [n], optional): flip / rot90 code of each
    sample (0 identity, 1 fliplr, 2 flipud, 3 both, 4/5/6 rot90 k=1/2/3), applied as an index remap."""
    if not img1.is_cuda:
        raise RuntimeError("gather_patches needs CUDA tensors (no CPU path)")
    assert img1.dtype == torch.float32 and img2.dtype == torch.float32
    img1, img2 = img1.contiguous(), img2.contiguous()
    H, W, C1 = img1.shape
    C2 = img2.shape[2]
    xy = xy.to(device=img1.device, dtype=torch.int32).contiguous()
    n = xy.shape[0]
    if validate:          # callers whose coordinates are valid by construction (MultiModalX) skip the read-back
        validate_xy(xy, H, W, P, center_mode)
    hsi = torch.empty(n, C1, P, P, dtype=torch.float32, device=img1.device)
    lid = torch.empty(n, C2, P, P, dtype=torch.float32, device=img1.device)
    labels, eb = None, 0
    if gt is not None:
        gt = gt.contiguous()
        eb = gt.element_size()
        if gt.dtype not in (torch.uint8, torch.int32, torch.int64):
            raise ValueError("gt must be uint8, int32 or int64")
        labels = torch.empty(n, dtype=torch.int64, device=img1.device)
    with torch.cuda.device(img1.device):
        if ops is not None:
            ops = ops.to(device=img1.device, dtype=torch.uint8).contiguous()
        _lib.check(_lib.lib().vc_gather_patches_f32(img1.data_ptr(), img2.data_ptr(), _ptr(gt), eb, H, W, C1, C2,
                                                    xy.data_ptr(), _ptr(ops), n, P, 1 if center_mode else 0, hsi.data_ptr(),
                                                    lid.data_ptr(), _ptr(labels), _stream()), "vc_gather_patches_f32")
    return hsi, lid, labels


def scene_index(xs, ys, first, count, W, C1, C2, P):
    dev = xs.device
    off1 = torch.empty(count, dtype=torch.int64, device=dev)
    off2 = torch.empty(count, dtype=torch.int64, device=dev)
    oidx = torch.empty(count, dtype=torch.int64, device=dev)
    xy = torch.empty(count, 2, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().vc_scene_index(xs.data_ptr(), ys.data_ptr(), xs.numel(), ys.numel(), first, count, W, C1,
                                             C2, P, off1.data_ptr(), off2.data_ptr(), oidx.data_ptr(), xy.data_ptr(),
                                             _stream()), "vc_scene_index")
    return off1, off2, oidx, xy


def pack_sps(x, S):
    """x f32 [n,C,P,P] (any strides) -> bf16 SPS buffer [S, rows, 8]."""
    n, C, P, _ = x.shape
    out = torch.empty(S, sps_rows(n, P), 8, dtype=torch.bfloat16, device=x.device)
    sb, sc, si, sj = x.stride()
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().vc_pack_sps(x.data_ptr(), sb, sc, si, sj, 0, n, C, P, out.data_ptr(), S, _stream()),
                   "vc_pack_sps")
    return out


def pack_sps_raster(img, patch_off, P, S):
    """raster f32 [H,W,C] + per-patch element offsets of the window corner -> SPS buffer."""
    H, W, C = img.shape
    n = patch_off.numel()
    out = torch.empty(S, sps_rows(n, P), 8, dtype=torch.bfloat16, device=img.device)
    with torch.cuda.device(img.device):
        _lib.check(_lib.lib().vc_pack_sps(img.data_ptr(), 0, 1, W * C, C, patch_off.data_ptr(), n, C, P,
                                          out.data_ptr(), S, _stream()), "vc_pack_sps")
    return out


def conv_sps(sps_in, w_packed, scale, bias, n, P, relu=True, impl=0, debug_flags=0, out=None, out_slice_off=0):
    """3x3 / 1x1 conv + affine (+ReLU) on SPS buffers; w_packed bf16 [ns][taps][S_in][ncta][8]."""
    ns, taps, s_in, ncta, _ = w_packed.shape
    n_out = ns * ncta
    assert sps_in.shape[0] == s_in and sps_in.shape[1] == sps_rows(n, P)
    if out is None:
        out = torch.zeros(n_out // 8, sps_in.shape[1], 8, dtype=torch.bfloat16, device=sps_in.device)
    with torch.cuda.device(sps_in.device):
        _lib.check(_lib.lib().vc_conv_sps(sps_in.data_ptr(), s_in, w_packed.data_ptr(), scale.data_ptr(),
                                          bias.data_ptr(), out.data_ptr(), out_slice_off, n_out, ns, n, P, taps,
                                          1 if relu else 0, impl, debug_flags, _stream()), "vc_conv_sps")
    return out


def tokens_forward(f_sps, tparams, n, P, K, out_index=None, logits=None, argmax_map=None):
    if logits is None:
        logits = torch.empty(n, K, dtype=torch.float32, device=f_sps.device)
    with torch.cuda.device(f_sps.device):
        _lib.check(_lib.lib().vc_tokens_forward(f_sps.data_ptr(), tparams.data_ptr(), n, P, K, logits.data_ptr(),
                                                _ptr(out_index), _ptr(argmax_map), _stream()), "vc_tokens_forward")
    return logits


def tokens_forward_tc(f_sps, tparams, n, P, K, out_index=None, logits=None, argmax_map=None, scratch=None):
    """Token stage on the tcgen05 kernel (P*P + 1 <= 128); same contract as tokens_forward."""
    L = _lib.lib()
    if logits is None:
        logits = torch.empty(n, K, dtype=torch.float32, device=f_sps.device)
    need = L.vc_tokens_tc_scratch_bytes(n)
    if scratch is None or scratch.numel() * scratch.element_size() < need:
        scratch = torch.empty(need, dtype=torch.uint8, device=f_sps.device)
    with torch.cuda.device(f_sps.device):
        _lib.check(L.vc_tokens_forward_tc(f_sps.data_ptr(), tparams.data_ptr(), n, P, K, logits.data_ptr(),
                                          _ptr(out_index), _ptr(argmax_map), scratch.data_ptr(),
                                          scratch.numel() * scratch.element_size(), _stream()), "vc_tokens_forward_tc")
    return logits


def wgrad_sps(a_sps, b_sps, n, P, taps, shift_on_a, out, M, N, sm, sn, st, bias_col=-1, out_bias=None,
              accumulate=False, workspace=None):
    """dW[tap][m][n] = sum_rows A[row+sa][m] * B[row+sb][n] over SPS buffers (tcgen05, MN-major
    operands); result scattered into ``out`` (fp32) with element strides (sm, sn, st)."""
    L = _lib.lib()
    SA, SB = a_sps.shape[0], b_sps.shape[0]
    need = L.vc_wgrad_workspace_bytes(SB, taps)
    if workspace is None or workspace.numel() < need:
        workspace = torch.empty(need, dtype=torch.uint8, device=a_sps.device)
    with torch.cuda.device(a_sps.device):
        _lib.check(L.vc_wgrad_sps(a_sps.data_ptr(), SA, b_sps.data_ptr(), SB, n, P, taps, 1 if shift_on_a else 0,
                                  workspace.data_ptr(), workspace.numel(), out.data_ptr(), M, N, sm, sn, st,
                                  bias_col, _ptr(out_bias), 1 if accumulate else 0, _stream()), "vc_wgrad_sps")
    return out

"""CPU emulation of what the CUDA kernels compute FROM THE PACKED BUFFERS (test helper).

It restates the kernels' data layouts (SPS activations, packed conv weights, token-stage
blob) with plain torch CPU ops, so the Python-side packing code can be validated against
the fp32 oracle without a GPU, and so a GPU mismatch can be attributed to a kernel rather
than to packing.  Rounding points (bf16 operands, fp32 accumulate) follow the kernels.
"""
from __future__ import annotations

import numpy as np
import torch


def halo(P):
    return (P + 2 + 7) // 8 * 8


def pp(P):
    return (P + 1) * (P + 1)


def tiles(n, P):
    return (n * pp(P) + 127) // 128


def rows(n, P):
    return 2 * halo(P) + 128 * tiles(n, P)


def bf16(x):
    return x.to(torch.bfloat16).to(torch.float32)


def row_index(n, P):
    """rows (with lead halo) of all valid pixels: [n, P, P] int64"""
    b = torch.arange(n).view(n, 1, 1)
    i = torch.arange(P).view(1, P, 1)
    j = torch.arange(P).view(1, 1, P)
    return halo(P) + b * pp(P) + i * (P + 1) + j


def pack_sps(x, S):
    """x f32 [n, C, P, P] -> SPS [S][RT][8] (values bf16-rounded, stored as f32)."""
    n, C, P, _ = x.shape
    flat = torch.zeros(rows(n, P), S * 8)
    R = row_index(n, P).reshape(-1)
    flat[R, :C] = bf16(x).permute(0, 2, 3, 1).reshape(-1, C)
    return flat.view(-1, S, 8).permute(1, 0, 2).contiguous()


def unpack_sps(sps, n, P, C):
    """SPS [S][RT][8] (any float dtype) -> [n, C, P, P] f32 (valid cells only)."""
    S, RT, _ = sps.shape
    flat = sps.float().permute(1, 0, 2).reshape(RT, S * 8)
    return flat[row_index(n, P).reshape(-1), :C].reshape(n, P, P, C).permute(0, 3, 1, 2).contiguous()


def conv_sps(sps_in, wpacked, scale, bias, n, P, relu=True):
    """sps_in [S_in][RT][8]; wpacked bf16 [ns][taps][S_in][ncta][8] -> [n_out/8][RT][8]."""
    S_in, RT, _ = sps_in.shape
    ns, taps, _, ncta, _ = wpacked.shape
    n_out = ns * ncta
    A = sps_in.permute(1, 0, 2).reshape(RT, S_in * 8)
    W = wpacked.float().permute(1, 0, 3, 2, 4).reshape(taps, n_out, S_in * 8)   # [tap][oc][c]
    H = halo(P)
    M = 128 * tiles(n, P)
    acc = torch.zeros(M, n_out)
    for tap in range(taps):
        shift = ((tap // 3 - 1) * (P + 1) + (tap % 3 - 1)) if taps == 9 else 0
        acc += A[H + shift:H + shift + M] @ W[tap].t()
    y = acc * scale.view(1, -1) + bias.view(1, -1)
    if relu:
        y = torch.relu(y)
    valid = torch.zeros(RT, dtype=torch.bool)
    valid[row_index(n, P).reshape(-1)] = True
    out = torch.zeros(RT, n_out)
    out[H:H + M] = bf16(y)
    out[~valid] = 0
    return out.view(RT, n_out // 8, 8).permute(1, 0, 2).contiguous()


def _get_bf16(blob, off, rows_, cols, pitch):
    v = blob[off:off + rows_ * pitch * 2].view(torch.bfloat16).view(rows_, pitch)
    return v[:, :cols].float()


def _get_f32(blob, off, count):
    return blob[off:off + count * 4].view(torch.float32).clone()


def _ln(x, g, b):
    m = x.mean(-1, keepdim=True)
    v = ((x - m) ** 2).mean(-1, keepdim=True)
    return (x - m) * torch.rsqrt(v + 1e-6) * g + b


def tokens_forward(f_sps, blob, layout, n, P, K):
    """Token stage from the fused feature buffer f_sps [8][RT][8] and the parameter blob."""
    D, T = 32, P * P + 1
    RT = f_sps.shape[1]
    F = f_sps.permute(1, 0, 2).reshape(RT, 64)
    feat = F[row_index(n, P).reshape(n, -1)]                       # [n, P*P, 64]
    wfus = _get_bf16(blob, layout["wfus"], D, 64, 72)
    x = feat @ wfus.t()
    x = torch.relu(x * _get_f32(blob, layout["fus_scale"], D) + _get_f32(blob, layout["fus_bias"], D))
    cls = _get_f32(blob, layout["cls"], D).view(1, 1, D).expand(n, 1, D)
    x = torch.cat([cls, x], 1) + _get_f32(blob, layout["pos"], T * D).view(1, T, D)
    for lo in layout["layers"]:
        y = bf16(_ln(x, _get_f32(blob, lo["ln1_g"], D), _get_f32(blob, lo["ln1_b"], D)))
        qkv = y @ _get_bf16(blob, lo["wqkv"], 3 * D, D, D + 8).t() + _get_f32(blob, lo["bqkv"], 3 * D)
        q, k, v = qkv.view(n, T, 3, 4, 8).permute(2, 0, 3, 1, 4)
        att = torch.softmax(bf16(q * 8 ** -0.5) @ bf16(k).transpose(-1, -2), -1)
        o = (bf16(att) @ bf16(v)).transpose(1, 2).reshape(n, T, D)
        x = x + bf16(o) @ _get_bf16(blob, lo["wproj"], D, D, D + 8).t() + _get_f32(blob, lo["bproj"], D)
        y = bf16(_ln(x, _get_f32(blob, lo["ln2_g"], D), _get_f32(blob, lo["ln2_b"], D)))
        h = y @ _get_bf16(blob, lo["wfc1"], 4 * D, D, D + 8).t() + _get_f32(blob, lo["bfc1"], 4 * D)
        h = bf16(torch.nn.functional.gelu(h))
        x = x + h @ _get_bf16(blob, lo["wfc2"], D, 4 * D, 4 * D + 8).t() + _get_f32(blob, lo["bfc2"], D)
    c = _ln(x[:, 0], _get_f32(blob, layout["lnf_g"], D), _get_f32(blob, layout["lnf_b"], D))
    return c @ _get_f32(blob, layout["whead"], K * D).view(K, D).t() + _get_f32(blob, layout["bhead"], K)


def model_forward(model, layout, hsi, lidar):
    """Whole eval forward from a vitcnn_b200.ViTCNN's packed parameters (CPU tensors)."""
    from vitcnn_b200.model import (HSI_PLANES, LIDAR_PLANES, fold_bn, pack_conv_weight, pack_tparams, slices_for,
                                   stem_plan)
    n, _, P, _ = hsi.shape
    outs = []
    for stem, x, planes in ((model.hsi_stem, hsi, HSI_PLANES), (model.lidar_stem, lidar, LIDAR_PLANES)):
        a = pack_sps(x, slices_for(x.shape[1]))
        for (s_in, n_out, ns), layer in zip(stem_plan(x.shape[1], planes), stem):
            w = pack_conv_weight(layer.conv.weight, s_in, n_out, ns)
            s, b = fold_bn(layer.conv, layer.bn, n_out)
            a = conv_sps(a, w, s, b, n, P)
        outs.append(a)
    f = torch.cat(outs, 0)
    blob = pack_tparams(model, layout)
    return tokens_forward(f, blob, layout, n, P, model.num_classes)


def py_tparams_layout(P, K):
    """Python restatement of csrc/vc_tparams.h::tlayout (used where the .so cannot be loaded
    and to cross-check vc_tparams_layout)."""
    D, HID, LAYERS = 32, 128, 2
    off = 0

    def take(nbytes):
        nonlocal off
        o = off
        off += (nbytes + 15) & ~15
        return o

    L = {"wfus": take(D * 72 * 2), "layers": []}
    for _ in range(LAYERS):
        L["layers"].append({"wqkv": take(3 * D * 40 * 2), "wproj": take(D * 40 * 2), "wfc1": take(HID * 40 * 2),
                            "wfc2": take(D * 136 * 2)})
    L["fus_scale"], L["fus_bias"], L["cls"] = take(D * 4), take(D * 4), take(D * 4)
    for l in range(LAYERS):
        d = L["layers"][l]
        d["ln1_g"], d["ln1_b"], d["bqkv"], d["bproj"] = take(D * 4), take(D * 4), take(3 * D * 4), take(D * 4)
        d["ln2_g"], d["ln2_b"], d["bfc1"], d["bfc2"] = take(D * 4), take(D * 4), take(HID * 4), take(D * 4)
    L["lnf_g"], L["lnf_b"] = take(D * 4), take(D * 4)
    L["whead"], L["bhead"] = take(K * D * 4), take(K * 4)
    L["pos"] = take((P * P + 1) * D * 4)
    L["total"] = off
    return L

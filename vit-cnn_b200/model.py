"""ViT-CNN hybrid as a ``torch.nn.Module`` whose forward runs in libvitcnn.so.

Interface = what the reference's factory builds for its ViT-style models
(model_utils.py:206-218): ``ViTCNN(n_bands, n_bands2, embed_dim, patch_size, patch_size_vit,
num_patches, nheads, num_layers, num_classes, dropout)`` and ``forward(hsi[B,C1,P,P],
lidar[B,C2,P,P]) -> logits f32 [B,K]`` as called at model_utils.py:921 / 1118 / 1144.
Architecture: SURVEY.md App. A ("R0").  The sub-modules below are parameter containers
only (they give the state_dict its timm-style keys and the default initialisation); no
torch operator computes anything on the forward / backward path.
"""
from __future__ import annotations

import ctypes
import math
import os

import numpy as np
import torch
import torch.nn as nn

from . import _lib

HSI_PLANES = (128, 64, 32)
LIDAR_PLANES = (8, 16, 32)
_W_BUDGET = 180_000   # bytes of shared memory one CTA may spend on resident conv weights
# Windows per launch group of scene inference.  Measured on B200, Houston scene (tools/chunk_sweep.py,
# profiles/r01_chunk_sweep.jsonl): 32768 -> 54.8 ms, 65536 -> 54.0, 131072 -> 53.5, 196608 / 262144 -> 52.8 (all maps
# bit-identical); the workspace is sized to min(chunk, windows of the call): 26 GB of the 180 GB at 131072.
SCENE_CHUNK = int(os.environ.get("VITCNN_CHUNK", "131072"))


# ----------------------------------------------------------------------------------------------
# parameter containers
# ----------------------------------------------------------------------------------------------
class ConvBn(nn.Module):
    def __init__(self, cin, cout, k):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, kernel_size=k, stride=1, padding=k // 2, bias=True)
        self.bn = nn.BatchNorm2d(cout)


class AttentionParams(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.qkv = nn.Linear(dim, 3 * dim, bias=True)
        self.proj = nn.Linear(dim, dim)


class MlpParams(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.fc2 = nn.Linear(hidden, dim)


class BlockParams(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.attn = AttentionParams(dim)
        self.norm2 = nn.LayerNorm(dim, eps=1e-6)
        self.mlp = MlpParams(dim, 4 * dim)


# ----------------------------------------------------------------------------------------------
# packing into kernel-ready form (pure tensor reshuffles; run once per weight version)
# ----------------------------------------------------------------------------------------------
def slices_for(c: int) -> int:
    return (c + 15) // 16 * 2


def choose_nsplit(s_in: int, n_out: int, taps: int) -> int:
    for ns in (1, 2, 4, 8):
        ncta = n_out // ns
        if n_out % ns == 0 and ncta % 16 == 0 and ncta <= 128 and taps * s_in * ncta * 16 <= _W_BUDGET:
            return ns
    raise ValueError(f"conv weights do not fit: S_in={s_in} n_out={n_out}")


def pack_conv_weight(w: torch.Tensor, s_in: int, n_out: int, nsplit: int) -> torch.Tensor:
    """[Cout, Cin, kh, kw] fp32 -> bf16 [nsplit][taps][S_in][n_out/nsplit][8] (zero padded)."""
    cout, cin, kh, kw = w.shape
    taps = kh * kw
    full = torch.zeros(n_out, s_in * 8, taps, dtype=torch.float32, device=w.device)
    full[:cout, :cin] = w.detach().float().reshape(cout, cin, taps)
    ncta = n_out // nsplit
    full = full.reshape(nsplit, ncta, s_in, 8, taps).permute(0, 4, 2, 1, 3)
    return full.contiguous().to(torch.bfloat16)


def fold_bn(conv: nn.Conv2d, bn: nn.BatchNorm2d, n_out: int):
    """Eval-mode BatchNorm + conv bias as a per-channel affine: y = conv_nobias * scale + bias."""
    scale = bn.weight.detach().float() / torch.sqrt(bn.running_var.float() + bn.eps)
    bias = (conv.bias.detach().float() - bn.running_mean.float()) * scale + bn.bias.detach().float()
    s = torch.zeros(n_out, dtype=torch.float32, device=scale.device)
    b = torch.zeros(n_out, dtype=torch.float32, device=scale.device)
    s[:scale.numel()] = scale
    b[:bias.numel()] = bias
    return s.contiguous(), b.contiguous()


def pack_tparams(model: "ViTCNN", layout: dict) -> torch.Tensor:
    """Token-stage parameter blob (csrc/vc_tparams.h) as a uint8 tensor on the model's device."""
    dev = model.cls_token.device
    blob = torch.zeros(layout["total"], dtype=torch.uint8, device=dev)

    def put_bf16(off, w, pitch):          # w [rows, cols] -> rows of `pitch` bf16
        rows, cols = w.shape
        view = blob[off:off + rows * pitch * 2].view(torch.bfloat16).view(rows, pitch)
        view[:, :cols] = w.detach().to(torch.bfloat16)

    def put_f32(off, v):
        v = v.detach().float().reshape(-1)
        blob[off:off + v.numel() * 4].view(torch.float32).copy_(v)

    D = model.embed_dim
    fconv, fbn = model.fusion.conv, model.fusion.bn
    put_bf16(layout["wfus"], fconv.weight.reshape(D, -1), fconv.weight.shape[1] + 8)
    fs, fb = fold_bn(fconv, fbn, D)
    put_f32(layout["fus_scale"], fs)
    put_f32(layout["fus_bias"], fb)
    put_f32(layout["cls"], model.cls_token)
    put_f32(layout["pos"], model.pos_embed)
    for blk, lo in zip(model.blocks, layout["layers"]):
        put_bf16(lo["wqkv"], blk.attn.qkv.weight, D + 8)
        put_bf16(lo["wproj"], blk.attn.proj.weight, D + 8)
        put_bf16(lo["wfc1"], blk.mlp.fc1.weight, D + 8)
        put_bf16(lo["wfc2"], blk.mlp.fc2.weight, 4 * D + 8)
        put_f32(lo["ln1_g"], blk.norm1.weight)
        put_f32(lo["ln1_b"], blk.norm1.bias)
        put_f32(lo["bqkv"], blk.attn.qkv.bias)
        put_f32(lo["bproj"], blk.attn.proj.bias)
        put_f32(lo["ln2_g"], blk.norm2.weight)
        put_f32(lo["ln2_b"], blk.norm2.bias)
        put_f32(lo["bfc1"], blk.mlp.fc1.bias)
        put_f32(lo["bfc2"], blk.mlp.fc2.bias)
    put_f32(layout["lnf_g"], model.norm.weight)
    put_f32(layout["lnf_b"], model.norm.bias)
    put_f32(layout["whead"], model.head.weight)
    put_f32(layout["bhead"], model.head.bias)
    return blob


def pack_lidar_blob(model: "ViTCNN") -> torch.Tensor:
    """Operands of the fused LiDAR-stem kernel (csrc/lidar_stem.cu; layout in include/vitcnn.h):
    bf16 weight rows of 24 elements (conv 1 / 2 with two taps per K=16 step, conv 3 tap-major),
    then the folded-BN affine of the three layers in fp32."""
    dev = model.cls_token.device
    l1, l2, l3 = model.lidar_stem
    c2 = model.n_bands2
    assert c2 <= 8
    rows = torch.zeros(5 * 8 + 5 * 16 + 9 * 32, 24, dtype=torch.float32, device=dev)
    w1 = l1.conv.weight.detach().float().reshape(8, c2, 9)
    w2 = l2.conv.weight.detach().float().reshape(16, 8, 9)
    w3 = l3.conv.weight.detach().float().reshape(32, 16, 9)
    for p in range(5):
        for half, tap in ((0, 2 * p), (1, 2 * p + 1)):
            if tap > 8:
                continue
            rows[p * 8:p * 8 + 8, 8 * half:8 * half + c2] = w1[:, :, tap]
            rows[40 + p * 16:40 + p * 16 + 16, 8 * half:8 * half + 8] = w2[:, :, tap]
    for tap in range(9):
        rows[120 + tap * 32:120 + tap * 32 + 32, :16] = w3[:, :, tap]
    aff = []
    for layer, n in ((l1, 8), (l2, 16), (l3, 32)):
        s, b = fold_bn(layer.conv, layer.bn, n)
        aff += [s, b]
    blob = torch.cat([rows.to(torch.bfloat16).reshape(-1).view(torch.uint8), torch.cat(aff).view(torch.uint8)])
    assert blob.numel() == _lib.lib().vc_lidar_blob_bytes()
    return blob.contiguous()


def stem_plan(c_in: int, planes) -> list:
    """(S_in, n_out_padded, nsplit) per conv of a 3-layer stem; channels are padded to 16."""
    plan, s_in = [], slices_for(c_in)
    for p in planes:
        n_out = (p + 15) // 16 * 16
        plan.append((s_in, n_out, choose_nsplit(s_in, n_out, 9)))
        s_in = n_out // 8
    return plan


# ----------------------------------------------------------------------------------------------
class ViTCNN(nn.Module):
    def __init__(self, n_bands, n_bands2, embed_dim=32, patch_size=11, patch_size_vit=1, num_patches=None,
                 nheads=4, num_layers=2, num_classes=16, dropout=0.01):
        super().__init__()
        num_patches = patch_size * patch_size if num_patches is None else num_patches
        if patch_size_vit != 1 or num_patches != patch_size * patch_size:
            raise ValueError("ViT-CNN tokenises one pixel per token: patch_size_vit=1, num_patches=P*P")
        if embed_dim != 32 or nheads != 4 or num_layers != 2:
            raise ValueError("kernels are specialised for embed_dim=32, nheads=4, num_layers=2 "
                             "(model_utils.py:206-218)")
        if not 1 <= patch_size <= 15:
            raise ValueError("patch_size must be in [1, 15]")
        if not 1 <= num_classes <= 64:
            raise ValueError("num_classes must be in [1, 64]")
        self.n_bands, self.n_bands2 = int(n_bands), int(n_bands2)
        self.embed_dim, self.patch_size = embed_dim, int(patch_size)
        self.nheads, self.num_layers, self.num_classes = nheads, num_layers, int(num_classes)
        self.dropout = float(dropout)
        a, b = HSI_PLANES, LIDAR_PLANES
        self.hsi_stem = nn.Sequential(ConvBn(n_bands, a[0], 3), ConvBn(a[0], a[1], 3), ConvBn(a[1], a[2], 3))
        self.lidar_stem = nn.Sequential(ConvBn(n_bands2, b[0], 3), ConvBn(b[0], b[1], 3), ConvBn(b[1], b[2], 3))
        self.fusion = ConvBn(a[2] + b[2], embed_dim, 1)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, num_patches + 1, embed_dim))
        self.blocks = nn.ModuleList([BlockParams(embed_dim) for _ in range(num_layers)])
        self.norm = nn.LayerNorm(embed_dim, eps=1e-6)
        self.head = nn.Linear(embed_dim, num_classes)
        self.reset_parameters()
        self._pack = None
        self._pack_key = None
        self._ws = None

    def reset_parameters(self):
        # vision_transformer.py:552-558,709-717; conv/BN init of the S2ENet idiom (App. A.2)
        nn.init.trunc_normal_(self.pos_embed, std=0.02)
        nn.init.normal_(self.cls_token, std=1e-6)
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.trunc_normal_(m.weight, std=0.02)
                nn.init.zeros_(m.bias)
            elif isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
                nn.init.zeros_(m.bias)
            elif isinstance(m, (nn.BatchNorm2d, nn.LayerNorm)):
                nn.init.ones_(m.weight)
                nn.init.zeros_(m.bias)

    # ---- kernel-ready parameters ------------------------------------------------------------
    def _version_key(self):
        ts = list(self.parameters()) + list(self.buffers())
        return (str(self.cls_token.device),) + tuple(t._version for t in ts) + tuple(t.data_ptr() for t in ts[:2])

    def pack_for_inference(self) -> dict:
        """Fold BN, convert and lay out all weights for the eval-mode kernels (cached per
        weight version)."""
        key = self._version_key()
        if self._pack is not None and self._pack_key == key:
            return self._pack
        P, K = self.patch_size, self.num_classes
        pk = {"keep": []}
        st = _lib.VcModel()
        st.C1, st.C2, st.P, st.K = self.n_bands, self.n_bands2, P, K
        st.S1, st.S2 = slices_for(self.n_bands), slices_for(self.n_bands2)
        for stem, c_in, planes, pre in ((self.hsi_stem, self.n_bands, HSI_PLANES, "h"),
                                        (self.lidar_stem, self.n_bands2, LIDAR_PLANES, "l")):
            for i, ((s_in, n_out, ns), layer) in enumerate(zip(stem_plan(c_in, planes), stem)):
                w = pack_conv_weight(layer.conv.weight, s_in, n_out, ns)
                s, b = fold_bn(layer.conv, layer.bn, n_out)
                pk["keep"] += [w, s, b]
                getattr(st, "w_" + pre)[i] = w.data_ptr()
                getattr(st, "scale_" + pre)[i] = s.data_ptr()
                getattr(st, "bias_" + pre)[i] = b.data_ptr()
                getattr(st, "nsplit_" + pre)[i] = ns
        # shared first conv of dense scene inference: conv-1 weights with the taps that leave a window dropped,
        # one copy per (row class, column class) of a window pixel (include/vitcnn.h: w_h1_border)
        w1 = self.hsi_stem[0].conv.weight.detach()
        s_in, n_out, ns = stem_plan(self.n_bands, HSI_PLANES)[0]
        border = []
        for cy in range(3):
            for cx in range(3):
                wv = w1.clone()
                if cy == 0:
                    wv[:, :, 0, :] = 0
                if cy == 2:
                    wv[:, :, 2, :] = 0
                if cx == 0:
                    wv[:, :, :, 0] = 0
                if cx == 2:
                    wv[:, :, :, 2] = 0
                border.append(pack_conv_weight(wv, s_in, n_out, ns).reshape(-1))
        wb = torch.cat(border).contiguous()
        pk["keep"].append(wb)
        st.w_h1_border = wb.data_ptr()
        blob = pack_tparams(self, _lib.tparams_layout(P, K))
        pk["keep"].append(blob)
        pk["tparams"] = blob
        st.tparams = blob.data_ptr()
        if self.n_bands2 <= 8:
            lb = pack_lidar_blob(self)
            pk["keep"].append(lb)
            st.lidar_blob = lb.data_ptr()
        pk["struct"] = st
        self._pack, self._pack_key = pk, key
        return pk

    def _workspace(self, n: int, device) -> torch.Tensor:
        need = _lib.lib().vc_workspace_bytes(n, self.patch_size, self.n_bands, self.n_bands2)
        if self._ws is None or self._ws.numel() < need or self._ws.device != device:
            self._ws = torch.empty(need, dtype=torch.uint8, device=device)
        return self._ws

    # ---- forward ------------------------------------------------------------------------------
    def forward(self, hsi: torch.Tensor, lidar: torch.Tensor) -> torch.Tensor:
        if not (hsi.is_cuda and lidar.is_cuda and self.cls_token.is_cuda):
            raise RuntimeError("ViTCNN runs on a CUDA device only (no CPU path): move the model and "
                               "its inputs to cuda")
        if hsi.dim() != 4 or lidar.dim() != 4 or hsi.shape[0] != lidar.shape[0]:
            raise ValueError("expected hsi [B,C1,P,P] and lidar [B,C2,P,P]")
        P = self.patch_size
        if tuple(hsi.shape[1:]) != (self.n_bands, P, P) or tuple(lidar.shape[1:]) != (self.n_bands2, P, P):
            raise ValueError(f"input shapes {tuple(hsi.shape)}, {tuple(lidar.shape)} do not match the model")
        if hsi.dtype != torch.float32:
            hsi = hsi.float()
        if lidar.dtype != torch.float32:
            lidar = lidar.float()
        if self.training:
            from .train import train_forward
            return train_forward(self, hsi, lidar)
        return self._forward_eval(hsi, lidar)

    def _forward_eval(self, hsi, lidar):
        L = _lib.lib()
        n = hsi.shape[0]
        logits = torch.empty(n, self.num_classes, dtype=torch.float32, device=hsi.device)
        if n == 0:
            return logits
        with torch.cuda.device(hsi.device):
            pk = self.pack_for_inference()
            ws = self._workspace(n, hsi.device)
            hs = (ctypes.c_int64 * 4)(*hsi.stride())
            ls = (ctypes.c_int64 * 4)(*lidar.stride())
            stream = torch.cuda.current_stream().cuda_stream
            _lib.check(L.vc_forward_patches(ctypes.byref(pk["struct"]), hsi.data_ptr(), hs, lidar.data_ptr(), ls, n,
                                            ws.data_ptr(), ws.numel(), logits.data_ptr(), stream),
                       "vc_forward_patches")
        return logits

    def _starts_on_device(self, starts, dev) -> torch.Tensor:
        """Window-start list as an int32 device tensor.  Cached by content: a pageable-host upload blocks the
        host until the stream has drained, which would serialise the sub-bands of predict_scene_host."""
        arr = np.ascontiguousarray(starts, dtype=np.int32)
        key = (str(dev), arr.tobytes())
        cache = self.__dict__.setdefault("_starts_cache", {})
        t = cache.get(key)
        if t is None:
            if len(cache) >= 256:
                cache.clear()
            t = cache[key] = torch.from_numpy(arr).to(dev)
        return t

    # ---- full-scene inference ---------------------------------------------------------------------
    @torch.no_grad()
    def predict_scene(self, img1: torch.Tensor, img2: torch.Tensor, stride: int = 1, chunk: int = SCENE_CHUNK,
                      window_range=None, logits_map=None, argmax_map=None, xs=None, ys=None):
        """Sliding-window inference over device-resident rasters img1 f32 [H,W,C1], img2 f32
        [H,W,C2] (the loop of test(), model_utils.py:1086-1129).  Returns (logits_map f32
        [H,W,K], argmax_map uint8 [H,W]); pixels no window is centred on stay 0.
        ``window_range=(first, count)`` restricts to a contiguous run of windows in the
        reference's row-major order (row-band sharding); ``xs`` / ``ys`` override the window
        start lists (a row band uploaded on its own keeps the scene's starts)."""
        from .utils import window_starts
        if self.training:
            raise RuntimeError("predict_scene is an eval-mode path: call model.eval() first")
        if not (img1.is_cuda and img2.is_cuda):
            raise RuntimeError("rasters must be CUDA tensors")
        img1 = img1.contiguous().float()
        img2 = img2.contiguous().float()
        H, W, C1 = img1.shape
        P, K = self.patch_size, self.num_classes
        if C1 != self.n_bands or img2.shape[2] != self.n_bands2 or tuple(img2.shape[:2]) != (H, W):
            raise ValueError("raster shapes do not match the model")
        dev = img1.device
        import numpy as np
        xs = window_starts(H, P, stride) if xs is None else np.ascontiguousarray(xs, dtype=np.int32)
        ys = window_starts(W, P, stride) if ys is None else np.ascontiguousarray(ys, dtype=np.int32)
        if len(xs) and (xs.min() < 0 or xs.max() + P > H) or len(ys) and (ys.min() < 0 or ys.max() + P > W):
            raise ValueError("window starts fall outside the raster")
        nx, ny = len(xs), len(ys)
        first, count = (0, nx * ny) if window_range is None else window_range
        if logits_map is None:
            logits_map = torch.zeros(H, W, K, dtype=torch.float32, device=dev)
        if argmax_map is None:
            argmax_map = torch.zeros(H, W, dtype=torch.uint8, device=dev)
        if count > 0:
            # hand the library only the raster rows this window range touches (a row band of a sharded scene):
            # its scene-level work (the shared first conv) then covers the band, not the whole raster
            r_first, r_last = first // ny, (first + count - 1) // ny
            y_lo, y_hi = int(xs[r_first:r_last + 1].min()), int(xs[r_first:r_last + 1].max()) + P
            xs = xs[r_first:r_last + 1] - y_lo
            first -= r_first * ny
            nx = len(xs)
            img1, img2 = img1[y_lo:y_hi], img2[y_lo:y_hi]
            lg_view, am_view = logits_map[y_lo:y_hi], argmax_map[y_lo:y_hi]
            H = y_hi - y_lo
            xs, ys = self._starts_on_device(xs, dev), self._starts_on_device(ys, dev)
            L = _lib.lib()
            with torch.cuda.device(dev):
                pk = self.pack_for_inference()
                # equal-sized chunks (no short tail launch): ceil(count / ceil(count / chunk))
                n_chunks = -(-count // int(chunk))
                chunk = -(-count // n_chunks)
                need = L.vc_scene_workspace_bytes(ctypes.byref(pk["struct"]), H, W, chunk)
                if self._ws is None or self._ws.numel() < need or self._ws.device != dev:
                    self._ws = torch.empty(need, dtype=torch.uint8, device=dev)
                ws = self._ws
                stream = torch.cuda.current_stream().cuda_stream
                _lib.check(L.vc_scene_infer(ctypes.byref(pk["struct"]), img1.data_ptr(), img2.data_ptr(), H, W,
                                            xs.data_ptr(), ys.data_ptr(), nx, ny, first, count, chunk, ws.data_ptr(),
                                            ws.numel(), lg_view.data_ptr(), am_view.data_ptr(), stream),
                           "vc_scene_infer")
        return logits_map, argmax_map

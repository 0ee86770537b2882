"""CPU oracle for the ViT-CNN hybrid ("R0" reconstruction) in plain fp32 PyTorch.

TEST INFRASTRUCTURE ONLY - see ``oracle/data_ref.py`` for the import rule.

PARITY: the TOKEN STAGE (cls / pos-embed, the two pre-norm blocks with MHSA and the GELU MLP, final
LayerNorm, cls pooling, head: everything after the fusion conv) is PINNED to the reference's own source:
``tests/golden/make_block_golden.py`` extracts ``Attention`` / ``Block`` / ``Mlp`` and
``VisionTransformer._pos_embed / forward_features / forward_head`` by AST from the vendored timm files
and runs them unchanged; ``tests/test_oracle_blocks_cpu.py`` checks this file against their outputs
(committed fixture ``tests/golden/block_golden.npz``; bit for bit against the live source when
``/root/reference`` is mounted).  The CNN STEM IS PARITY UNPINNED: the reference ships no ``ViT-CNN`` /
``FICNN_VIT`` source, no checkpoint, no test and no golden vector for it, so that
half is a reconstruction from the surviving evidence, not a transcription:

* constructor contract and training recipe  - model_utils.py:206-218
* conv_bn_relu idiom (Conv2d 3x3 pad 1 bias -> BatchNorm2d -> ReLU), planes
  (128,64,32) / (8,16,32), 1x1 fusion on the channel concat - S2ENet bytecode
  (SURVEY.md App. A.2), model/Multimodality_Mamba/Mutimodality_Mamba7.py:1119-1153
* per-pixel tokens (patch_size_vit = 1, row-major) -
  model/compare_method/vit/timm/layers/patch_embed.py:65,89
* cls token, pos-embed, pre-norm block, MHSA, MLP, final norm, cls pooling, head -
  model/compare_method/vit/timm/models/vision_transformer.py:57-105,123-166,463,
  598-629,680-701 and model/compare_method/vit/timm/layers/mlp.py:13-47
* initialisation - vision_transformer.py:552-558,709-717; S2ENet bytecode.

The CUDA module loads this module's ``state_dict`` unchanged (same keys).
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

STEM_HSI_PLANES = (128, 64, 32)
STEM_LIDAR_PLANES = (8, 16, 32)
LN_EPS = 1e-6      # vision_transformer.py:463
BN_EPS = 1e-5      # nn.BatchNorm2d default
BN_MOMENTUM = 0.1


class ConvBnRelu(nn.Module):
    def __init__(self, cin, cout, k=3, pad=1):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, kernel_size=k, stride=1, padding=pad, bias=True)
        self.bn = nn.BatchNorm2d(cout)

    def forward(self, x):
        return F.relu(self.bn(self.conv(x)))


class Attention(nn.Module):
    """vision_transformer.py:57-105 (qkv_bias=True, no qk-norm, non-fused branch)."""

    def __init__(self, dim, num_heads, attn_drop=0.0, proj_drop=0.0):
        super().__init__()
        self.num_heads = num_heads
        self.head_dim = dim // num_heads
        self.scale = self.head_dim ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=True)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)

    def forward(self, x):
        B, N, C = x.shape
        qkv = self.qkv(x).reshape(B, N, 3, self.num_heads, self.head_dim).permute(2, 0, 3, 1, 4)
        q, k, v = qkv.unbind(0)
        attn = (q * self.scale) @ k.transpose(-2, -1)
        attn = self.attn_drop(attn.softmax(dim=-1))
        x = (attn @ v).transpose(1, 2).reshape(B, N, C)
        return self.proj_drop(self.proj(x))


class Mlp(nn.Module):
    """mlp.py:13-47."""

    def __init__(self, dim, hidden, drop=0.0):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.fc2 = nn.Linear(hidden, dim)
        self.drop = nn.Dropout(drop)

    def forward(self, x):
        return self.drop(self.fc2(self.drop(F.gelu(self.fc1(x)))))


class Block(nn.Module):
    """vision_transformer.py:123-166 (no LayerScale / DropPath)."""

    def __init__(self, dim, num_heads, mlp_ratio=4.0, drop=0.0):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=LN_EPS)
        self.attn = Attention(dim, num_heads, attn_drop=0.0, proj_drop=drop)
        self.norm2 = nn.LayerNorm(dim, eps=LN_EPS)
        self.mlp = Mlp(dim, int(dim * mlp_ratio), drop=drop)

    def forward(self, x):
        x = x + self.attn(self.norm1(x))
        x = x + self.mlp(self.norm2(x))
        return x


class ViTCNNRef(nn.Module):
    """fp32 reference of the hybrid; signature from model_utils.py:213."""

    def __init__(self, n_bands, n_bands2, embed_dim=32, patch_size=11, patch_size_vit=1,
                 num_patches=None, nheads=4, num_layers=2, num_classes=16, dropout=0.01):
        super().__init__()
        assert patch_size_vit == 1, "one token per pixel (model_utils.py:213)"
        num_patches = patch_size * patch_size if num_patches is None else num_patches
        assert num_patches == patch_size * patch_size
        assert embed_dim == STEM_HSI_PLANES[-1]
        self.n_bands, self.n_bands2 = n_bands, n_bands2
        self.embed_dim, self.patch_size = embed_dim, patch_size
        self.nheads, self.num_layers, self.num_classes = nheads, num_layers, num_classes
        a, b = STEM_HSI_PLANES, STEM_LIDAR_PLANES
        self.hsi_stem = nn.Sequential(ConvBnRelu(n_bands, a[0]), ConvBnRelu(a[0], a[1]), ConvBnRelu(a[1], a[2]))
        self.lidar_stem = nn.Sequential(ConvBnRelu(n_bands2, b[0]), ConvBnRelu(b[0], b[1]), ConvBnRelu(b[1], b[2]))
        self.fusion = ConvBnRelu(a[2] + b[2], embed_dim, k=1, pad=0)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, num_patches + 1, embed_dim))
        self.pos_drop = nn.Dropout(dropout)
        self.blocks = nn.ModuleList([Block(embed_dim, nheads, 4.0, dropout) for _ in range(num_layers)])
        self.norm = nn.LayerNorm(embed_dim, eps=LN_EPS)
        self.head = nn.Linear(embed_dim, num_classes)
        self.reset_parameters()

    def reset_parameters(self):
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

    def embed_tokens(self, x):
        """vision_transformer.py:598-629 (_pos_embed, class token, no_embed_class False): x [B, P*P, D]."""
        x = torch.cat([self.cls_token.expand(x.shape[0], -1, -1), x], dim=1)
        return self.pos_drop(x + self.pos_embed)

    def tokens(self, hsi, lidar):
        h = self.hsi_stem(hsi)
        l = self.lidar_stem(lidar)
        f = self.fusion(torch.cat([h, l], dim=1))           # [B, D, P, P]
        return self.embed_tokens(f.flatten(2).transpose(1, 2))     # [B, P*P, D], row-major pixels (patch_embed.py:65,89)

    def forward_tokens(self, x):
        """vision_transformer.py:682-702 (forward_features after the embedding, forward_head with cls pooling)."""
        for blk in self.blocks:
            x = blk(x)
        x = self.norm(x)
        return self.head(x[:, 0])

    def forward(self, hsi, lidar):
        return self.forward_tokens(self.tokens(hsi, lidar))


def randomize_bn_stats(model: nn.Module, seed: int = 1) -> None:
    """Give every BatchNorm non-trivial affine parameters and running statistics
    so eval-mode parity exercises the BN fold (fresh BN is the identity)."""
    g = torch.Generator().manual_seed(seed)
    for m in model.modules():
        if isinstance(m, nn.BatchNorm2d):
            with torch.no_grad():
                m.weight.copy_(0.75 + 0.5 * torch.rand(m.weight.shape, generator=g))
                m.bias.copy_(0.2 * torch.randn(m.bias.shape, generator=g))
                m.running_mean.copy_(0.1 * torch.randn(m.running_mean.shape, generator=g))
                m.running_var.copy_(0.5 + torch.rand(m.running_var.shape, generator=g))


def forward_flops(C1, C2, P, K, D=32, heads=4, layers=2, ratio=4):
    """Algorithmic forward FLOPs (2*MAC) per sample, SURVEY.md App. D."""
    px, T, d = P * P, P * P + 1, D // heads
    a, b = STEM_HSI_PLANES, STEM_LIDAR_PLANES
    stem = 2 * px * 9 * (C1 * a[0] + a[0] * a[1] + a[1] * a[2])
    lid = 2 * px * 9 * (C2 * b[0] + b[0] * b[1] + b[1] * b[2])
    fus = 2 * px * (a[2] + b[2]) * D
    qkv = layers * 2 * T * D * 3 * D
    proj = layers * 2 * T * D * D
    attn = 2 * layers * heads * 2 * T * T * d
    mlp = layers * 2 * 2 * T * D * ratio * D
    head = 2 * D * K
    return dict(stem=stem, lidar=lid, fusion=fus, qkv=qkv, proj=proj, attn=attn, mlp=mlp,
                head=head, total=stem + lid + fus + qkv + proj + attn + mlp + head)

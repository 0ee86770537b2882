"""Pin the transformer half of the model oracle to the reference's own source.

The reference vendors timm's ViT (model/compare_method/vit/timm/...), the only surviving trace of the
ViT half of "ViT-CNN" (SURVEY.md F5).  The package cannot be imported as a whole (timm.utils, _features,
_hub ... are absent), but the classes the token stage is made of are plain torch code.  This script
extracts them BY AST from the files where they lie and executes them unchanged:

  vision_transformer.py:57-105   class Attention          (non-fused branch: use_fused_attn() -> False)
  vision_transformer.py:108-120  class LayerScale         (unused: init_values=None)
  vision_transformer.py:123-166  class Block
  layers/mlp.py:13-47            class Mlp
  vision_transformer.py:598-629  VisionTransformer._pos_embed      (cls concat, + pos_embed, dropout)
  vision_transformer.py:682-692  VisionTransformer.forward_features
  vision_transformer.py:694-702  VisionTransformer.forward_head    (cls pooling, head)

  model/compare_method/FusAtNet.py:9-17  class ConvUnit  (Conv2d 3x3 pad 1 bias -> BatchNorm2d -> ReLU: the
                                 conv_bn_relu idiom every stem layer of the reconstruction is an instance of;
                                 which planes / how many layers the stem has stays a reconstruction)

The three methods are bound to a minimal shell module carrying exactly the attributes they read, built
the way VisionTransformer.__init__ builds them for (embed_dim 32, depth 2, heads 4, mlp_ratio 4,
qkv_bias True, class_token, global_pool 'token', LayerNorm eps 1e-6 (:463), patch embedding = the CNN's
fused map, so patch_embed is the identity here).

Run in the dev container (needs /root/reference):  python tests/golden/make_block_golden.py
Output: tests/golden/block_golden.npz (seeded weights, inputs, and the extracted code's outputs).
tests/test_oracle_blocks_cpu.py checks oracle/model_ref.py against it (and live, bit for bit, when
/root/reference is mounted)."""
from __future__ import annotations

import ast
import os
import sys
from functools import partial
from typing import Optional  # noqa: F401  (used by the extracted source)

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F  # noqa: F401

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
TIMM = os.path.join(os.environ.get("VITCNN_REFERENCE_ROOT", "/root/reference"), "model", "compare_method", "vit", "timm")
VIT_PY = os.path.join(TIMM, "models", "vision_transformer.py")
MLP_PY = os.path.join(TIMM, "layers", "mlp.py")
FUS_PY = os.path.join(os.environ.get("VITCNN_REFERENCE_ROOT", "/root/reference"), "model", "compare_method", "FusAtNet.py")


def available() -> bool:
    return os.path.isfile(VIT_PY) and os.path.isfile(MLP_PY) and os.path.isfile(FUS_PY)


def _segments(path, class_names=(), methods_of=None, method_names=()):
    src = open(path).read()
    tree = ast.parse(src)
    out = {}
    for node in tree.body:
        if isinstance(node, ast.ClassDef) and node.name in class_names:
            out[node.name] = ast.get_source_segment(src, node)
        if isinstance(node, ast.ClassDef) and node.name == methods_of:
            for sub in node.body:
                if isinstance(sub, ast.FunctionDef) and sub.name in method_names:
                    out[sub.name] = ast.get_source_segment(src, sub)
    return out


def extract():
    """Returns (Attention, Block, Mlp, TokenStage) built from the reference's source text."""
    import textwrap
    from torch.jit import Final  # noqa: F401
    ns = {"nn": nn, "torch": torch, "F": F, "Final": Final, "Optional": Optional, "partial": partial,
          "use_fused_attn": lambda: False,            # timm.layers.use_fused_attn: the non-fused branch is the spec
          "to_2tuple": lambda v: tuple(v) if isinstance(v, (tuple, list)) else (v, v),   # timm.layers.helpers
          "DropPath": nn.Identity,                    # drop_path = 0 in Block.__init__ -> never instantiated
          "checkpoint_seq": None, "resample_abs_pos_embed": None}
    mlp = _segments(MLP_PY, class_names=("Mlp",))
    exec(compile(mlp["Mlp"], MLP_PY, "exec"), ns)
    vit = _segments(VIT_PY, class_names=("Attention", "LayerScale", "Block"), methods_of="VisionTransformer",
                    method_names=("_pos_embed", "forward_features", "forward_head", "forward"))
    for name in ("Attention", "LayerScale", "Block"):
        exec(compile(vit[name], VIT_PY, "exec"), ns)
    methods = {}
    for name in ("_pos_embed", "forward_features", "forward_head", "forward"):
        exec(compile(textwrap.dedent(vit[name]), VIT_PY, "exec"), ns, methods)
    Block = ns["Block"]

    class TokenStage(nn.Module):
        """The attributes VisionTransformer.__init__ (vision_transformer.py:394-560) sets for this configuration."""

        def __init__(self, num_tokens, num_classes, embed_dim=32, depth=2, num_heads=4, mlp_ratio=4.0, drop_rate=0.0):
            super().__init__()
            norm_layer = partial(nn.LayerNorm, eps=1e-6)                                  # :463
            self.num_prefix_tokens, self.no_embed_class, self.dynamic_img_size = 1, False, False
            self.global_pool, self.grad_checkpointing = "token", False
            self.patch_embed = nn.Identity()
            self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
            self.reg_token = None
            self.pos_embed = nn.Parameter(torch.randn(1, num_tokens, embed_dim) * .02)
            self.pos_drop = nn.Dropout(p=drop_rate)
            self.patch_drop, self.norm_pre = nn.Identity(), nn.Identity()
            self.blocks = nn.Sequential(*[Block(dim=embed_dim, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=True,
                                                proj_drop=drop_rate, norm_layer=norm_layer, act_layer=nn.GELU)
                                          for _ in range(depth)])
            self.norm = norm_layer(embed_dim)
            self.attn_pool = None
            self.fc_norm, self.head_drop = nn.Identity(), nn.Dropout(drop_rate)
            self.head = nn.Linear(embed_dim, num_classes)

    for name, fn in methods.items():
        setattr(TokenStage, name, fn)
    return ns["Attention"], Block, ns["Mlp"], TokenStage


def extract_conv_unit():
    """FusAtNet.ConvUnit from the reference's source text."""
    ns = {"nn": nn, "torch": torch}
    exec(compile(_segments(FUS_PY, class_names=("ConvUnit",))["ConvUnit"], FUS_PY, "exec"), ns)
    return ns["ConvUnit"]


def seeded_conv_unit(ConvUnit, cin, cout, seed):
    torch.manual_seed(seed)
    m = ConvUnit(cin, cout)
    g = torch.Generator().manual_seed(seed + 50)
    with torch.no_grad():
        m.bn.weight.copy_(0.75 + 0.5 * torch.rand(cout, generator=g))
        m.bn.bias.copy_(0.2 * torch.randn(cout, generator=g))
        m.bn.running_mean.copy_(0.1 * torch.randn(cout, generator=g))
        m.bn.running_var.copy_(0.5 + torch.rand(cout, generator=g))
        m.conv.bias.copy_(0.1 * torch.randn(cout, generator=g))
    return m


CONV_CASES = (("c144", 144, 16, 11, 2), ("c1", 1, 8, 7, 3), ("c64", 64, 32, 9, 2))


def seeded_token_stage(TokenStage, P, K, seed):
    torch.manual_seed(seed)
    m = TokenStage(P * P + 1, K)
    with torch.no_grad():
        for p in m.parameters():           # non-trivial norms / biases / cls token; sharper attention than std 0.02
            p.copy_(torch.randn_like(p) * (0.3 if p.dim() > 1 else 0.2) + (1.0 if p.dim() == 1 and p.numel() == 32 else 0.0))
    return m.eval()


CASES = (("p11", 11, 16, 3, 0), ("p7", 7, 12, 5, 1), ("p5", 5, 4, 2, 2))


def main():
    assert available(), TIMM
    Attention, Block, Mlp, TokenStage = extract()
    out = {}
    for name, P, K, B, seed in CASES:
        m = seeded_token_stage(TokenStage, P, K, seed)
        g = torch.Generator().manual_seed(100 + seed)
        x = torch.randn(B, P * P, 32, generator=g)
        with torch.no_grad():
            t0 = m._pos_embed(x)
            b0 = m.blocks[0](t0)
            a0 = m.blocks[0].attn(m.blocks[0].norm1(t0))
            h0 = m.blocks[0].mlp(m.blocks[0].norm2(t0))
            logits = m(x)
        out[f"{name}_cfg"] = np.array([P, K, B, seed], dtype=np.int64)
        out[f"{name}_x"] = x.numpy()
        for k, v in m.state_dict().items():
            out[f"{name}_sd_{k}"] = v.numpy()
        for k, v in (("tokens", t0), ("block0", b0), ("attn0", a0), ("mlp0", h0), ("logits", logits)):
            out[f"{name}_{k}"] = v.numpy()
    ConvUnit = extract_conv_unit()
    for name, cin, cout, P, B in CONV_CASES:
        m = seeded_conv_unit(ConvUnit, cin, cout, cin)
        x = torch.rand(B, cin, P, P, generator=torch.Generator().manual_seed(cin))
        out[f"{name}_cfg"] = np.array([cin, cout, P, B], dtype=np.int64)
        out[f"{name}_x"] = x.numpy()
        for k, v in m.state_dict().items():
            out[f"{name}_sd_{k}"] = v.numpy().copy()          # the training forward below updates the running stats
        with torch.no_grad():
            out[f"{name}_eval"] = m.eval()(x).numpy()
        out[f"{name}_train"] = m.train()(x).detach().numpy()          # batch statistics; running stats updated
        out[f"{name}_running_mean_after"] = m.bn.running_mean.numpy().copy()
        out[f"{name}_running_var_after"] = m.bn.running_var.numpy().copy()
    path = os.path.join(HERE, "block_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items() if k.endswith("logits")})


if __name__ == "__main__":
    sys.path.insert(0, ROOT)
    main()

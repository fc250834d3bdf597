"""Factorised audio-visual fusion block (mirror of reference models/fusion_blocks.py:33-59,
216-289).  Only ``factorized_mmi`` is on the hot path: every reference YAML selects it
(configs/deepavfusion.yaml:28); the ``token`` / ``dense_mmi`` variants are out of scope
(SURVEY.md 2.1 #2).

Parameter names / shapes are the reference's:
    norm1_mm, norm1_aud, norm1_img, norm2 : LayerNorm(dim)            (eps 1e-5)
    attn.attn_v / attn.attn_a             : q [D,D], kv [2D,D], proj [D,D]
    attn.q [D*r, D], attn.k [D*r, 2D], attn.v [D, 2D], attn.proj [D, D]
    mlp.fc1 [D*mlp_ratio, D], mlp.fc2
"""
from __future__ import annotations

from types import SimpleNamespace

import torch
from torch import nn

from .. import functional as Fn
from .layers import Mlp, _no_dropout


class CrossAttention(nn.Module):
    """fusion_blocks.py:33-44 parameter holder (q / kv / proj)."""

    def __init__(self, dim, num_heads=8, qkv_bias=False):
        super().__init__()
        self.num_heads = num_heads
        self.scale = (dim // num_heads) ** -0.5
        self.q = nn.Linear(dim, dim, bias=qkv_bias)
        self.kv = nn.Linear(dim, dim * 2, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim)

    def _ns(self):
        return SimpleNamespace(q_w=self.q.weight, q_b=self.q.bias, kv_w=self.kv.weight, kv_b=self.kv.bias,
                               proj_w=self.proj.weight, proj_b=self.proj.bias)


class CrossAttention_FactorizedAVInteractions(nn.Module):
    """fusion_blocks.py:216-233 parameter holder."""

    def __init__(self, dim, num_heads=8, qkv_bias=False, attn_drop=0.0, proj_drop=0.0, dim_ratio=1.0, fusion_tkns=(8, 4, 4)):
        super().__init__()
        _no_dropout(attn_drop=attn_drop, proj_drop=proj_drop)
        self.num_heads = num_heads
        self.scale = (dim // num_heads) ** -0.5            # NOT a function of dim_ratio (:220-222)
        self.dim = int(dim * dim_ratio)
        self.fusion_tkns = tuple(fusion_tkns)
        self.attn_v = CrossAttention(dim, num_heads=num_heads, qkv_bias=qkv_bias)
        self.attn_a = CrossAttention(dim, num_heads=num_heads, qkv_bias=qkv_bias)
        self.q = nn.Linear(dim, self.dim, bias=qkv_bias)
        self.k = nn.Linear(dim * 2, self.dim, bias=qkv_bias)
        self.v = nn.Linear(dim * 2, dim, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim)


class FusionBlock_FactorizedAVInteractions(nn.Module):
    """fusion_blocks.py:266-289.  forward(xmm, xv, xa) -> updated fusion tokens; the residual is taken
    from the NORMED fusion tokens (:281-283) exactly like the reference."""

    def __init__(self, dim, num_heads, attn_ratio=0.25, mlp_ratio=4.0, qkv_bias=False, fusion_tkns=(8, 4, 4),
                 drop=0.0, attn_drop=0.0, drop_path=0.0, act_layer=nn.GELU, norm_layer=nn.LayerNorm):
        super().__init__()
        _no_dropout(drop=drop, attn_drop=attn_drop)
        assert act_layer is nn.GELU
        self.drop_path = float(drop_path)          # fusion_blocks.py:276: one DropPath module, two independent draws (:283,:288)
        self.norm1_mm = norm_layer(dim)
        self.norm1_aud = norm_layer(dim)
        self.norm1_img = norm_layer(dim)
        self.attn = CrossAttention_FactorizedAVInteractions(
            dim, num_heads=num_heads, qkv_bias=qkv_bias, dim_ratio=attn_ratio, fusion_tkns=fusion_tkns)
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(dim, int(dim * mlp_ratio))
        self._ns = None

    def _bind(self, store):
        at = self.attn
        a = SimpleNamespace(store=store, heads=at.num_heads, tkns=at.fusion_tkns, eps=self.norm1_mm.eps,
                            n_mm_w=self.norm1_mm.weight, n_mm_b=self.norm1_mm.bias,
                            n_img_w=self.norm1_img.weight, n_img_b=self.norm1_img.bias,
                            n_aud_w=self.norm1_aud.weight, n_aud_b=self.norm1_aud.bias,
                            attn_v=at.attn_v._ns(), attn_a=at.attn_a._ns(),
                            q_w=at.q.weight, q_b=at.q.bias, k_w=at.k.weight, k_b=at.k.bias,
                            v_w=at.v.weight, v_b=at.v.bias, proj_w=at.proj.weight, proj_b=at.proj.bias)
        f = SimpleNamespace(store=store, eps=self.norm2.eps, norm_w=self.norm2.weight, norm_b=self.norm2.bias,
                            fc1_w=self.mlp.fc1.weight, fc1_b=self.mlp.fc1.bias,
                            fc2_w=self.mlp.fc2.weight, fc2_b=self.mlp.fc2.bias)
        self._ns = (a, f)

    def forward(self, xmm: torch.Tensor, xv: torch.Tensor, xa: torch.Tensor, return_attention: bool = False) -> torch.Tensor:
        if return_attention:
            raise NotImplementedError("return_attention is a visualisation path, not on the training hot path")
        a, f = self._ns
        d1 = d2 = None
        if self.drop_path > 0.0 and self.training:
            d1 = Fn.droppath_scale(xmm.shape[0], self.drop_path, xmm.device)
            d2 = Fn.droppath_scale(xmm.shape[0], self.drop_path, xmm.device)
        xmm = Fn.FusionAttnFn.apply(xmm, xv, xa, a.n_mm_w, a, d1)
        return Fn.MlpBranchFn.apply(xmm, f.norm_w, f, d2)

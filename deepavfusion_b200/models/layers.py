"""Parameter containers with timm 0.9.2's sub-module names (so ``state_dict`` keys match the
reference / the public MAE ViT-B checkpoint, vits.py:64-80) whose arithmetic is the fused
sm_100a kernels of ``deepavfusion_b200.functional`` rather than ATen.

``nn.Linear`` / ``nn.LayerNorm`` / ``nn.Conv2d`` are used purely as named parameter holders; their
own ``forward`` is never called on the hot path.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Optional

import torch
from torch import nn

from .. import functional as Fn
from ..params import ParamStore


# Tests flip this to exercise the host orchestration on CPU with kernels.* monkey-patched by
# torch emulations (tests/cpu_kernels.py).  The product never changes it.
_REQUIRE_CUDA = True


def _no_dropout(**kw):
    for k, v in kw.items():
        if v:
            raise NotImplementedError(
                f"{k}={v}: dropout is not on the hot path (every reference config uses 0); only 0 is implemented")


class PatchEmbed(nn.Module):
    """timm PatchEmbed(img_size, patch_size, in_chans, embed_dim): Conv2d k = s = patch, lowered to
    an im2col-free GEMM over the KEPT patches only (functional.PatchEmbedFn)."""

    def __init__(self, img_size, patch_size, in_chans, embed_dim):
        super().__init__()
        img_size = (img_size, img_size) if isinstance(img_size, int) else tuple(img_size)
        patch_size = (patch_size, patch_size) if isinstance(patch_size, int) else tuple(patch_size)
        assert patch_size[0] == patch_size[1], "square patches only"
        self.img_size, self.patch_size = img_size, patch_size
        self.grid_size = (img_size[0] // patch_size[0], img_size[1] // patch_size[1])
        self.num_patches = self.grid_size[0] * self.grid_size[1]
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)


class Attention(nn.Module):
    def __init__(self, dim, num_heads):
        super().__init__()
        self.num_heads = num_heads
        self.qkv = nn.Linear(dim, dim * 3, bias=True)
        self.proj = nn.Linear(dim, dim)


class Mlp(nn.Module):
    def __init__(self, in_features, hidden_features):
        super().__init__()
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.fc2 = nn.Linear(hidden_features, in_features)


class Block(nn.Module):
    """timm Block(dim, num_heads, mlp_ratio, qkv_bias=True, norm_layer=...): pre-LN attention + MLP.

    ``forward(x, prefix=None)``: ``prefix`` rows (fusion tokens) take part as keys / values only and
    the block returns the updated ``x`` rows -- exactly what deepavfusion.py:104-105 keeps of
    ``blk(cat(x_fusion, x))``."""

    def __init__(self, dim, num_heads, mlp_ratio=4.0, qkv_bias=True, norm_layer=nn.LayerNorm,
                 drop_path=0.0, attn_drop=0.0, proj_drop=0.0):
        super().__init__()
        assert qkv_bias, "qkv_bias=False is not used by the reference"
        _no_dropout(attn_drop=attn_drop, proj_drop=proj_drop)
        self.drop_path = float(drop_path)          # timm Block.drop_path1 / drop_path2 (same rate, independent draws)
        self.norm1 = norm_layer(dim)
        self.attn = Attention(dim, num_heads)
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(dim, int(dim * mlp_ratio))
        self._ns = None

    def _bind(self, store: ParamStore):
        a = SimpleNamespace(store=store, heads=self.attn.num_heads, eps=self.norm1.eps,
                            norm_w=self.norm1.weight, norm_b=self.norm1.bias,
                            qkv_w=self.attn.qkv.weight, qkv_b=self.attn.qkv.bias,
                            proj_w=self.attn.proj.weight, proj_b=self.attn.proj.bias)
        f = SimpleNamespace(store=store, eps=self.norm2.eps, norm_w=self.norm2.weight, norm_b=self.norm2.bias,
                            fc1_w=self.mlp.fc1.weight, fc1_b=self.mlp.fc1.bias,
                            fc2_w=self.mlp.fc2.weight, fc2_b=self.mlp.fc2.bias)
        self._ns = (a, f)

    def forward(self, x: torch.Tensor, prefix: Optional[torch.Tensor] = None) -> torch.Tensor:
        a, f = self._ns
        d1 = d2 = None
        if self.drop_path > 0.0 and self.training:      # stochastic depth: one per-sample draw per residual branch
            d1 = Fn.droppath_scale(x.shape[0], self.drop_path, x.device)
            d2 = Fn.droppath_scale(x.shape[0], self.drop_path, x.device)
        x = Fn.AttnBranchFn.apply(prefix, x, a.norm_w, a, d1)
        return Fn.MlpBranchFn.apply(x, f.norm_w, f, d2)


class FinalNorm(nn.LayerNorm):
    """nn.LayerNorm holder evaluated by the fused kernel (f32 in / f32 out)."""

    def _bind(self, store: ParamStore):
        self._ns = SimpleNamespace(store=store, eps=self.eps, norm_w=self.weight, norm_b=self.bias)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return Fn.LayerNormFn.apply(x, self.weight, self._ns)


def bind_all(root: nn.Module, store: ParamStore) -> None:
    for mod in root.modules():
        if hasattr(mod, "_bind"):
            mod._bind(store)
        mod.__dict__["_davf_store"] = store


def ensure_store(root: nn.Module) -> ParamStore:
    """Create (or re-create after ``.to()`` / ``.cuda()``) the flat parameter store of ``root`` and
    refresh the bf16 weight shadows if any parameter changed since the last forward."""
    store: Optional[ParamStore] = root.__dict__.get("_davf_store")
    if store is not None and store._depth > 0:
        return store                                   # called from inside an enclosing forward
    p0 = next(root.parameters())
    if _REQUIRE_CUDA and not p0.is_cuda:
        raise RuntimeError("deepavfusion_b200 runs on CUDA (sm_100a) only: move the model to the GPU first "
                           "(there is no CPU path; the CPU oracle under oracle/ is test infrastructure)")
    if store is None or not store.covers(root):
        store = ParamStore(root)
        bind_all(root, store)
    else:
        store.refresh_lowp()
    return store

"""Uni-modal ViT backbone (mirror of reference models/vits.py:16-141; ViT-B is the only size on
the BASELINE configs, the other constructors are kept because they are one-liners).

Same constructor, attributes (``embed_dim``, ``patch_embed.{grid_size,patch_size,num_patches}``,
``pos_embed``, ``blocks``, ``norm``), ``state_dict`` keys and ``load_checkpoint`` contract
(``strict=True`` against the public MAE ViT-B checkpoint, vits.py:64-80) as the reference.
"""
from __future__ import annotations

from functools import partial
from types import SimpleNamespace

import torch
from torch import nn

from .. import functional as Fn
from ..util.pos_embed import get_2d_sincos_pos_embed
from .layers import Block, FinalNorm, PatchEmbed, ensure_store

PRETRAINED_WEIGHTS = {
    "vit_base_audiomae_as2m": ("assets/models/vitbase_audiomae_as2m.pth", ""),
    "vit_base_mae_in1k": ("https://dl.fbaipublicfiles.com/mae/pretrain/mae_pretrain_vit_base.pth", ""),
    "vit_large_mae_in1k": ("https://dl.fbaipublicfiles.com/mae/pretrain/mae_pretrain_vit_large.pth", ""),
    "vit_huge_mae_in1k": ("https://dl.fbaipublicfiles.com/mae/pretrain/mae_pretrain_vit_huge.pth", ""),
}


def _xavier_linear_(mod: nn.Module) -> None:
    """vits.py:54-62: xavier-uniform Linear weights, zero biases, unit LayerNorm."""
    if isinstance(mod, nn.Linear):
        nn.init.xavier_uniform_(mod.weight)
        if mod.bias is not None:
            nn.init.constant_(mod.bias, 0)
    elif isinstance(mod, nn.LayerNorm):
        nn.init.constant_(mod.bias, 0)
        nn.init.constant_(mod.weight, 1.0)


class ViT(nn.Module):
    def __init__(self, input_size=224, patch_size=16, in_chans=3, embed_dim=1024, depth=24, num_heads=16,
                 mlp_ratio=4.0, norm_layer=nn.LayerNorm, use_cls_token=False, drop_path=0.0, attn_drop=0.0, drop=0.0):
        super().__init__()
        if use_cls_token:
            raise NotImplementedError("use_cls_token=True is never selected by the reference's DeepAVFusion (deepavfusion.py:20-21)")
        self.embed_dim = embed_dim
        self.patch_embed = PatchEmbed(input_size, patch_size, in_chans, embed_dim)
        self.pos_embed = nn.Parameter(torch.zeros(1, self.patch_embed.num_patches, embed_dim), requires_grad=False)
        self.cls_token = None
        eps = norm_layer.keywords.get("eps", 1e-5) if isinstance(norm_layer, partial) else 1e-5
        ln = partial(nn.LayerNorm, eps=eps)
        self.blocks = nn.ModuleList([
            Block(embed_dim, num_heads, mlp_ratio, qkv_bias=True, norm_layer=ln, drop_path=drop_path, attn_drop=attn_drop, proj_drop=drop)
            for _ in range(depth)])
        self.norm = FinalNorm(embed_dim, eps=eps)
        self.initialize_weights()

    def initialize_weights(self):
        pe = get_2d_sincos_pos_embed(self.pos_embed.shape[-1], self.patch_embed.grid_size, cls_token=False)
        self.pos_embed.data.copy_(torch.from_numpy(pe).float().unsqueeze(0))
        w = self.patch_embed.proj.weight.data
        nn.init.xavier_uniform_(w.view([w.shape[0], -1]))          # like nn.Linear, vits.py:43-44
        self.apply(_xavier_linear_)

    def load_checkpoint(self, ckpt_fn, prefix="", skip_keys_prefix=("decoder", "mask_token")):
        """vits.py:64-80: load an MAE-style checkpoint with strict=True (pos_embed kept as built)."""
        try:
            ckpt = torch.load(ckpt_fn, map_location="cpu")
        except Exception:
            ckpt = torch.hub.load_state_dict_from_url(url=ckpt_fn, map_location="cpu")
        if "state_dict" in ckpt:
            ckpt = ckpt["state_dict"]
        elif "model" in ckpt:
            ckpt = ckpt["model"]
        ckpt = {k[len(prefix):]: v for k, v in ckpt.items() if k.startswith(prefix)}
        ckpt = {k: v for k, v in ckpt.items() if not k.startswith(skip_keys_prefix)}
        if self.cls_token is None and "cls_token" in ckpt:
            del ckpt["cls_token"]
        ckpt["pos_embed"] = self.state_dict()["pos_embed"]
        self.load_state_dict(ckpt, strict=True)

    def params_layer_ids(self):
        """vits.py:82-89 (the (None, 0) entry for the absent cls_token is kept: lr_sched.py:32 builds
        a dict from these pairs)."""
        ids = [(p, 0) for p in self.patch_embed.parameters()]
        ids.append((self.cls_token, 0))
        for i, blk in enumerate(self.blocks):
            ids.extend([(p, i + 1) for p in blk.parameters()])
        ids.extend([(p, len(self.blocks) + 1) for p in self.norm.parameters()])
        return ids

    def _bind(self, store):
        self._pe_ns = SimpleNamespace(store=store, patch=self.patch_embed.patch_size[0],
                                      weight=self.patch_embed.proj.weight, bias=self.patch_embed.proj.bias,
                                      pos_embed=self.pos_embed)

    def prepare_patch_tokens(self, x, ids_keep=None):
        """vits.py:91-107: patch-embed + pos-embed (+ gather of the kept tokens) in one GEMM whose A
        rows are only the kept patches and whose epilogue adds bias + pos_embed[ids_keep]."""
        ensure_store(self)
        ns = self._pe_ns
        return Fn.PatchEmbedFn.apply(x.float().contiguous(), ids_keep, ns.weight, ns)

    def forward(self, x, ids_keep=None):
        x = self.prepare_patch_tokens(x, ids_keep=ids_keep)
        for blk in self.blocks:
            x = blk(x)
        return self.norm(x)


def vit_small_patch16(pretrained=False, **kwargs):
    assert pretrained is False or pretrained is None or pretrained == ""
    return ViT(patch_size=16, embed_dim=384, depth=12, num_heads=6, mlp_ratio=4, norm_layer=partial(nn.LayerNorm, eps=1e-6), **kwargs)


def vit_base_patch16(pretrained=False, **kwargs):
    model = ViT(patch_size=16, embed_dim=768, depth=12, num_heads=12, mlp_ratio=4, norm_layer=partial(nn.LayerNorm, eps=1e-6), **kwargs)
    if pretrained is not None and pretrained != "" and pretrained is not False:
        assert pretrained in {"vit_base_mae_in1k", "vit_base_audiomae_as2m"}
        url, prefix = PRETRAINED_WEIGHTS[pretrained]
        model.load_checkpoint(url, prefix=prefix)
    return model


def vit_large_patch16(pretrained=False, **kwargs):
    model = ViT(patch_size=16, embed_dim=1024, depth=24, num_heads=16, mlp_ratio=4, norm_layer=partial(nn.LayerNorm, eps=1e-6), **kwargs)
    if pretrained is not None and pretrained != "" and pretrained is not False:
        assert pretrained in {"vit_large_mae_in1k"}
        url, prefix = PRETRAINED_WEIGHTS[pretrained]
        model.load_checkpoint(url, prefix=prefix)
    return model


vit_small = vit_small_patch16
vit_base = vit_base_patch16
vit_large = vit_large_patch16

"""CPU oracle: a functional (no nn.Module) restatement of the reference's pre-training path.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Plain PyTorch on CPU tensors, fp32 by
default; ``amp=True`` emulates CUDA bf16 autocast (Linear / conv / matmul operands and results
in bf16 with fp32 accumulation, LayerNorm / softmax / loss in fp32, fp32 residual stream).

Every function cites the reference file:line it follows (paths under /root/reference).  The
per-modality transformer block, attention, MLP and patch-embed arithmetic lives in the
un-vendored third-party dependency ``timm==0.9.2`` (requirements.yml:21); its published
algorithm is restated here (timm/models/vision_transformer.py::{Attention,Block},
timm/layers/{patch_embed,mlp}.py) and anchored on the reference's call sites
(vits.py:27,32-34; avmae.py:53-55,83-85; fusion_blocks.py:278).

State is a flat ``dict[str, Tensor]`` with exactly the reference ``AVMAE.state_dict()`` keys.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# --------------------------------------------------------------------------------------
# configuration
# --------------------------------------------------------------------------------------
@dataclass
class OracleConfig:
    """Shapes of the ViT-B early-fusion MAE (configs/deepavfusion.yaml:12-35; vits.py:130-141)."""

    image_size: Tuple[int, int] = (224, 224)
    audio_size: Tuple[int, int] = (128, 192)
    patch: int = 16
    image_chans: int = 3
    audio_chans: int = 1
    dim: int = 768
    depth: int = 12
    heads: int = 12
    mlp_ratio: float = 4.0
    enc_eps: float = 1e-6            # vits.py:133  partial(nn.LayerNorm, eps=1e-6)
    fusion_tkns: Tuple[int, int, int] = (16, 8, 8)   # (nmm, nv, na) deepavfusion.yaml:30-32
    fusion_layers: object = "all"    # deepavfusion.py:38-46
    fusion_attn_ratio: float = 0.25
    fusion_mlp_ratio: float = 1.0
    fusion_heads: int = 12
    fusion_eps: float = 1e-5         # deepavfusion.py:50,52  plain nn.LayerNorm
    dec_dim: int = 512               # avmae.py:15
    dec_depth: int = 8
    dec_heads: int = 16
    dec_mlp_ratio: float = 4.0
    dec_eps: float = 1e-5            # avmae.py:15 norm_layer=nn.LayerNorm
    image_mask_ratio: float = 0.75
    audio_mask_ratio: float = 0.8
    image_norm_loss: bool = True
    audio_norm_loss: bool = True

    @property
    def image_grid(self):
        return (self.image_size[0] // self.patch, self.image_size[1] // self.patch)

    @property
    def audio_grid(self):
        return (self.audio_size[0] // self.patch, self.audio_size[1] // self.patch)

    def fusion_layer_set(self):
        """deepavfusion.py:37-46."""
        fl = self.fusion_layers
        if fl == "all":
            return set(range(self.depth))
        if fl == "none":
            return set()
        if isinstance(fl, int):
            return {fl}
        return {int(x) for x in str(fl).split("-")}


# --------------------------------------------------------------------------------------
# positional embedding  (util/pos_embed.py:42-90)
# --------------------------------------------------------------------------------------
def _sincos_1d(embed_dim: int, pos: np.ndarray) -> np.ndarray:
    """util/pos_embed.py:72-90: [sin | cos] of pos * 10000^(-i/(D/2))."""
    omega = np.arange(embed_dim // 2, dtype=np.float32)
    omega /= embed_dim / 2.0
    omega = 1.0 / 10000 ** omega
    out = np.einsum("m,d->md", pos.reshape(-1), omega)
    return np.concatenate([np.sin(out), np.cos(out)], axis=1)


def sincos_2d(embed_dim: int, grid_size: Tuple[int, int]) -> np.ndarray:
    """util/pos_embed.py:42-69.  Note the quirk kept from the reference: ``np.meshgrid(grid_w,
    grid_h)`` puts w first, the result is reshaped to (2,1,gH,gW), and the *first* half of the
    channels encodes grid[0] (the w coordinate)."""
    gh, gw = grid_size
    grid_h = np.arange(gh, dtype=np.float32)
    grid_w = np.arange(gw, dtype=np.float32)
    grid = np.stack(np.meshgrid(grid_w, grid_h), axis=0).reshape(2, 1, gh, gw)
    emb_a = _sincos_1d(embed_dim // 2, grid[0])
    emb_b = _sincos_1d(embed_dim // 2, grid[1])
    return np.concatenate([emb_a, emb_b], axis=1)


# --------------------------------------------------------------------------------------
# state dict construction (key set + shapes == reference AVMAE.state_dict())
# --------------------------------------------------------------------------------------
def _block_shapes(prefix: str, dim: int, hidden: int) -> Dict[str, Tuple[int, ...]]:
    """timm Block parameter names (vision_transformer.py::Block / Attention / Mlp)."""
    return {
        f"{prefix}.norm1.weight": (dim,), f"{prefix}.norm1.bias": (dim,),
        f"{prefix}.attn.qkv.weight": (3 * dim, dim), f"{prefix}.attn.qkv.bias": (3 * dim,),
        f"{prefix}.attn.proj.weight": (dim, dim), f"{prefix}.attn.proj.bias": (dim,),
        f"{prefix}.norm2.weight": (dim,), f"{prefix}.norm2.bias": (dim,),
        f"{prefix}.mlp.fc1.weight": (hidden, dim), f"{prefix}.mlp.fc1.bias": (hidden,),
        f"{prefix}.mlp.fc2.weight": (dim, hidden), f"{prefix}.mlp.fc2.bias": (dim,),
    }


def state_shapes(cfg: OracleConfig) -> Dict[str, Tuple[int, ...]]:
    """Ordered key -> shape map, in the registration order of the reference modules
    (avmae.py:26-89 registers encoder first, then audio decoder, then image decoder;
    deepavfusion.py:20-52; vits.py:26-35; fusion_blocks.py:216-278)."""
    D, p = cfg.dim, cfg.patch
    s: Dict[str, Tuple[int, ...]] = {}
    for mod, chans, grid in (("image", cfg.image_chans, cfg.image_grid), ("audio", cfg.audio_chans, cfg.audio_grid)):
        pre = f"encoder.{mod}"
        s[f"{pre}.pos_embed"] = (1, grid[0] * grid[1], D)
        s[f"{pre}.patch_embed.proj.weight"] = (D, chans, p, p)
        s[f"{pre}.patch_embed.proj.bias"] = (D,)
        for i in range(cfg.depth):
            s.update(_block_shapes(f"{pre}.blocks.{i}", D, int(D * cfg.mlp_ratio)))
        s[f"{pre}.norm.weight"] = (D,)
        s[f"{pre}.norm.bias"] = (D,)
        if mod == "image":
            pass
    # nn.Module registers parameters before sub-modules in state_dict(): fusion_tokens is a
    # direct Parameter of DeepAVFusion so it precedes encoder.image.* in the real ordering.
    # Ordering is irrelevant for parity (dict lookup); kept close to the reference for reading.
    s = {"encoder.fusion_tokens": (1, sum(cfg.fusion_tkns), D), **s}
    qk = int(D * cfg.fusion_attn_ratio)
    hid = int(D * cfg.fusion_mlp_ratio)
    for i in sorted(cfg.fusion_layer_set()):
        pre = f"encoder.fusion_blocks.{i}"
        for n in ("norm1_mm", "norm1_aud", "norm1_img"):
            s[f"{pre}.{n}.weight"] = (D,)
            s[f"{pre}.{n}.bias"] = (D,)
        for a in ("attn_v", "attn_a"):
            s[f"{pre}.attn.{a}.q.weight"] = (D, D); s[f"{pre}.attn.{a}.q.bias"] = (D,)
            s[f"{pre}.attn.{a}.kv.weight"] = (2 * D, D); s[f"{pre}.attn.{a}.kv.bias"] = (2 * D,)
            s[f"{pre}.attn.{a}.proj.weight"] = (D, D); s[f"{pre}.attn.{a}.proj.bias"] = (D,)
        s[f"{pre}.attn.q.weight"] = (qk, D); s[f"{pre}.attn.q.bias"] = (qk,)
        s[f"{pre}.attn.k.weight"] = (qk, 2 * D); s[f"{pre}.attn.k.bias"] = (qk,)
        s[f"{pre}.attn.v.weight"] = (D, 2 * D); s[f"{pre}.attn.v.bias"] = (D,)
        s[f"{pre}.attn.proj.weight"] = (D, D); s[f"{pre}.attn.proj.bias"] = (D,)
        s[f"{pre}.norm2.weight"] = (D,); s[f"{pre}.norm2.bias"] = (D,)
        s[f"{pre}.mlp.fc1.weight"] = (hid, D); s[f"{pre}.mlp.fc1.bias"] = (hid,)
        s[f"{pre}.mlp.fc2.weight"] = (D, hid); s[f"{pre}.mlp.fc2.bias"] = (D,)
    s["encoder.fusion_norm.weight"] = (D,)
    s["encoder.fusion_norm.bias"] = (D,)
    Dd = cfg.dec_dim
    for mod, chans, grid in (("audio", cfg.audio_chans, cfg.audio_grid), ("image", cfg.image_chans, cfg.image_grid)):
        s[f"{mod}_decoder_mask_token"] = (1, 1, Dd)
        s[f"{mod}_decoder_pos_embed"] = (1, grid[0] * grid[1], Dd)
        s[f"{mod}_decoder_embed.weight"] = (Dd, D); s[f"{mod}_decoder_embed.bias"] = (Dd,)
        for i in range(cfg.dec_depth):
            s.update(_block_shapes(f"{mod}_decoder_blocks.{i}", Dd, int(Dd * cfg.dec_mlp_ratio)))
        s[f"{mod}_decoder_norm.weight"] = (Dd,); s[f"{mod}_decoder_norm.bias"] = (Dd,)
        s[f"{mod}_decoder_pred.weight"] = (p * p * chans, Dd); s[f"{mod}_decoder_pred.bias"] = (p * p * chans,)
    return s


FROZEN_KEYS = ("encoder.image.pos_embed", "encoder.audio.pos_embed")   # vits.py:29 requires_grad=False


def build_state(cfg: OracleConfig, seed: int = 0, dtype=torch.float32) -> Dict[str, Tensor]:
    """Random-init state with the reference's *distributions* (xavier-uniform Linear weights,
    N(0,.02) tokens, sin-cos pos-embeds; vits.py:38-62, deepavfusion.py:56-68, avmae.py:92-118)
    but NOT its RNG order -- parity tests always copy one state dict into both sides.  Biases
    and LayerNorm affine terms get small random values instead of 0/1 so that tests exercise
    them."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, Tensor] = {}
    for k, shp in state_shapes(cfg).items():
        if k.endswith("pos_embed"):
            grid = cfg.image_grid if "image" in k else cfg.audio_grid
            t = torch.from_numpy(sincos_2d(shp[-1], grid)).float().unsqueeze(0)
        elif k.endswith("fusion_tokens") or k.endswith("mask_token"):
            t = torch.randn(shp, generator=g) * 0.02
        elif "norm" in k and k.endswith(".weight"):
            t = 1.0 + 0.1 * torch.randn(shp, generator=g)
        elif k.endswith(".bias"):
            t = 0.02 * torch.randn(shp, generator=g)
        else:  # Linear / conv weight: xavier uniform on the (out, fan_in) view
            fan_out = shp[0]
            fan_in = int(np.prod(shp[1:]))
            a = math.sqrt(6.0 / (fan_in + fan_out))
            t = (torch.rand(shp, generator=g) * 2 - 1) * a
        sd[k] = t.to(dtype)
    return sd


# --------------------------------------------------------------------------------------
# arithmetic helpers (autocast emulation)
# --------------------------------------------------------------------------------------
class _Prec:
    """amp=False: everything fp32.  amp=True: CUDA bf16-autocast semantics -- ``linear`` /
    ``matmul`` cast operands to bf16 and return bf16; ``layer_norm`` / ``softmax`` return fp32."""

    def __init__(self, amp: bool, sdpa: bool = False):
        self.amp = amp
        self.sdpa = sdpa      # ViT-block attention through F.scaled_dot_product_attention (timm 0.9.2 ``fused_attn``):
                              # only bench.py's eager-bf16-on-GPU leg sets it; every parity use keeps the explicit form

    def linear(self, x: Tensor, w: Tensor, b: Optional[Tensor]) -> Tensor:
        if self.amp:
            return F.linear(x.bfloat16(), w.bfloat16(), None if b is None else b.bfloat16())
        return F.linear(x, w, b)

    def matmul(self, a: Tensor, b: Tensor) -> Tensor:
        if self.amp:
            return torch.matmul(a.bfloat16(), b.bfloat16())
        return torch.matmul(a, b)

    def layer_norm(self, x: Tensor, w: Tensor, b: Tensor, eps: float) -> Tensor:
        return F.layer_norm(x.float(), (x.shape[-1],), w.float(), b.float(), eps)

    def softmax(self, x: Tensor) -> Tensor:
        return torch.softmax(x.float(), dim=-1)


# --------------------------------------------------------------------------------------
# a1: random masking  (avmae.py:120-142)
# --------------------------------------------------------------------------------------
def random_masking(noise: Tensor, mask_ratio: float):
    """avmae.py:120-142 with the noise passed in (the reference draws ``torch.rand(N, L)`` at
    :127).  Ties are broken lower-index-first (``stable=True``); the reference's own tie order is
    unspecified (``torch.argsort`` without ``stable``), see SURVEY.md section 7.3.
    Returns (ids_keep i64 (N,len_keep), mask f32 (N,L) 1=removed, ids_restore i64 (N,L))."""
    N, L = noise.shape
    ids_shuffle = torch.argsort(noise, dim=1, stable=True)
    ids_restore = torch.argsort(ids_shuffle, dim=1, stable=True)
    len_keep = int(L * (1 - mask_ratio))                      # :132
    ids_keep = ids_shuffle[:, :len_keep]
    mask = torch.ones(N, L, dtype=torch.float32, device=noise.device)
    mask[:, :len_keep] = 0
    mask = torch.gather(mask, 1, ids_restore)                 # :140
    return ids_keep, mask, ids_restore


# --------------------------------------------------------------------------------------
# a2: patch embed + pos + gather  (vits.py:91-107; timm PatchEmbed)
# --------------------------------------------------------------------------------------
def prepare_patch_tokens(P: _Prec, sd, pre: str, x: Tensor, ids_keep: Optional[Tensor], patch: int) -> Tensor:
    w, b = sd[f"{pre}.patch_embed.proj.weight"], sd[f"{pre}.patch_embed.proj.bias"]
    if P.amp:
        t = F.conv2d(x.bfloat16(), w.bfloat16(), b.bfloat16(), stride=patch)
    else:
        t = F.conv2d(x, w, b, stride=patch)                   # timm PatchEmbed.proj, k = s = 16
    t = t.flatten(2).transpose(1, 2)                          # (B, L, D)
    t = t + sd[f"{pre}.pos_embed"]                            # vits.py:96 (fp32 add promotes)
    if ids_keep is not None:                                  # vits.py:99-100
        t = t.gather(1, ids_keep.unsqueeze(-1).expand(-1, -1, t.shape[-1]))
    return t


# --------------------------------------------------------------------------------------
# a3: timm Block (pre-LN attention + MLP)  -- used by encoder (vits.py:32-34) and decoders
# --------------------------------------------------------------------------------------
def _attention(P: _Prec, sd, pre: str, x: Tensor, heads: int) -> Tensor:
    """timm 0.9.2 Attention.forward: fused qkv Linear, softmax(q k^T / sqrt(hd)) v, proj."""
    B, N, C = x.shape
    hd = C // heads
    qkv = P.linear(x, sd[f"{pre}.qkv.weight"], sd[f"{pre}.qkv.bias"])
    qkv = qkv.reshape(B, N, 3, heads, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    if P.sdpa:
        o = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B, N, C)
        return P.linear(o, sd[f"{pre}.proj.weight"], sd[f"{pre}.proj.bias"])
    s = P.matmul(q, k.transpose(-2, -1)).float() * hd ** -0.5
    a = P.softmax(s)
    o = P.matmul(a.to(v.dtype) if P.amp else a, v)
    o = o.transpose(1, 2).reshape(B, N, C)
    return P.linear(o, sd[f"{pre}.proj.weight"], sd[f"{pre}.proj.bias"])


def _mlp(P: _Prec, sd, pre: str, x: Tensor) -> Tensor:
    """timm Mlp: fc1 -> exact (erf) GELU -> fc2."""
    h = P.linear(x, sd[f"{pre}.fc1.weight"], sd[f"{pre}.fc1.bias"])
    return P.linear(F.gelu(h), sd[f"{pre}.fc2.weight"], sd[f"{pre}.fc2.bias"])


def _dp(drops, x: Tensor) -> Tensor:
    """timm DropPath (layers/drop.py): x * keep_mask / keep_prob with one value per sample.  ``drops`` is None
    (Identity) or an iterator of per-sample scale vectors [B], consumed in the reference's call order."""
    if drops is None:
        return x
    s = next(drops)
    return x * s.view(-1, *([1] * (x.dim() - 1))).to(x.dtype)


def vit_block(P: _Prec, sd, pre: str, x: Tensor, heads: int, eps: float, drops=None) -> Tensor:
    """timm Block.forward with ls* = Identity:
    x = x + drop_path1(attn(norm1(x))); x = x + drop_path2(mlp(norm2(x)))."""
    x = x + _dp(drops, _attention(P, sd, f"{pre}.attn", P.layer_norm(x, sd[f"{pre}.norm1.weight"], sd[f"{pre}.norm1.bias"], eps), heads))
    x = x + _dp(drops, _mlp(P, sd, f"{pre}.mlp", P.layer_norm(x, sd[f"{pre}.norm2.weight"], sd[f"{pre}.norm2.bias"], eps)))
    return x


# --------------------------------------------------------------------------------------
# a4: factorized fusion block  (fusion_blocks.py:33-59, 216-289)
# --------------------------------------------------------------------------------------
def _cross_attention(P: _Prec, sd, pre: str, x1: Tensor, x2: Tensor, heads: int) -> Tensor:
    """fusion_blocks.py:46-59 (CrossAttention.forward), explicit softmax path."""
    B, N1, C = x1.shape
    N2 = x2.shape[1]
    hd = C // heads
    q = P.linear(x1, sd[f"{pre}.q.weight"], sd[f"{pre}.q.bias"]).reshape(B, N1, heads, hd).permute(0, 2, 1, 3)
    kv = P.linear(x2, sd[f"{pre}.kv.weight"], sd[f"{pre}.kv.bias"]).reshape(B, N2, 2, heads, hd).permute(2, 0, 3, 1, 4)
    k, v = kv[0], kv[1]
    a = P.softmax(P.matmul(q, k.transpose(-2, -1)) * hd ** -0.5)          # :53-54
    o = P.matmul(a, v).transpose(1, 2).reshape(B, N1, C)                   # :57
    return P.linear(o, sd[f"{pre}.proj.weight"], sd[f"{pre}.proj.bias"])


def _factorized_attention(P: _Prec, sd, pre: str, xmm: Tensor, xv: Tensor, xa: Tensor,
                          tkns: Sequence[int], heads: int) -> Tensor:
    """fusion_blocks.py:235-263.  Scale is (dim/heads)^-0.5 irrespective of attn_ratio (:220-222)."""
    B, _, C = xmm.shape
    nmm, nv, na = tkns
    scale = (C // heads) ** -0.5
    xmm2, xmm_v, xmm_a = xmm.split((nmm, nv, na), dim=1)                   # :240
    xmm_v = _cross_attention(P, sd, f"{pre}.attn_v", xmm_v, xv, heads)    # :241
    xmm_a = _cross_attention(P, sd, f"{pre}.attn_a", xmm_a, xa, heads)    # :242
    xva = torch.cat((xmm_v.unsqueeze(2).expand(-1, -1, na, -1),
                     xmm_a.unsqueeze(1).expand(-1, nv, -1, -1)), dim=3).flatten(1, 2)   # :245-248
    q = P.linear(xmm2, sd[f"{pre}.q.weight"], sd[f"{pre}.q.bias"]).reshape(B, nmm, heads, -1).permute(0, 2, 1, 3)
    k = P.linear(xva, sd[f"{pre}.k.weight"], sd[f"{pre}.k.bias"]).reshape(B, nv * na, heads, -1).permute(0, 2, 1, 3)
    v = P.linear(xva, sd[f"{pre}.v.weight"], sd[f"{pre}.v.bias"]).reshape(B, nv * na, heads, -1).permute(0, 2, 1, 3)
    a = P.softmax(P.matmul(q, k.transpose(-2, -1)) * scale)               # :254-255
    o = P.matmul(a, v).transpose(1, 2).flatten(2)                          # :258
    o = P.linear(o, sd[f"{pre}.proj.weight"], sd[f"{pre}.proj.bias"])     # :259
    return torch.cat((o, xmm_v, xmm_a), dim=1)                             # :262


def fusion_block(P: _Prec, sd, pre: str, xmm: Tensor, xv: Tensor, xa: Tensor, cfg: OracleConfig, drops=None) -> Tensor:
    """fusion_blocks.py:280-289.  The residual is taken from the *normed* fusion tokens (:281-283)."""
    e = cfg.fusion_eps
    xmm = P.layer_norm(xmm, sd[f"{pre}.norm1_mm.weight"], sd[f"{pre}.norm1_mm.bias"], e)
    xv = P.layer_norm(xv, sd[f"{pre}.norm1_img.weight"], sd[f"{pre}.norm1_img.bias"], e)
    xa = P.layer_norm(xa, sd[f"{pre}.norm1_aud.weight"], sd[f"{pre}.norm1_aud.bias"], e)
    xmm = xmm + _dp(drops, _factorized_attention(P, sd, f"{pre}.attn", xmm, xv, xa, cfg.fusion_tkns, cfg.fusion_heads))   # :283
    xmm = xmm + _dp(drops, _mlp(P, sd, f"{pre}.mlp", P.layer_norm(xmm, sd[f"{pre}.norm2.weight"], sd[f"{pre}.norm2.bias"], e)))  # :288
    return xmm


# --------------------------------------------------------------------------------------
# a5: encoder  (deepavfusion.py:88-118)
# --------------------------------------------------------------------------------------
def encoder_forward(P: _Prec, sd, cfg: OracleConfig, image: Tensor, audio: Tensor,
                    image_ids_keep: Optional[Tensor] = None, audio_ids_keep: Optional[Tensor] = None,
                    return_embs: bool = False, drops=None):
    """``drops``: None, or an iterator over the DropPath scale vectors in the reference's call order -- per layer:
    image block (attention, MLP), audio block (attention, MLP), fusion block (attention, MLP)."""
    B = image.shape[0]
    x_i = prepare_patch_tokens(P, sd, "encoder.image", image, image_ids_keep, cfg.patch)   # :92
    x_a = prepare_patch_tokens(P, sd, "encoder.audio", audio, audio_ids_keep, cfg.patch)   # :93
    x_f = sd["encoder.fusion_tokens"].expand(B, -1, -1)                                    # :97
    nI, nA, nF = x_i.shape[1], x_a.shape[1], x_f.shape[1]
    fl = cfg.fusion_layer_set()
    embs = []
    for l in range(cfg.depth):                                                             # :99
        bi, ba = f"encoder.image.blocks.{l}", f"encoder.audio.blocks.{l}"
        if l not in fl:                                                                    # :100-102
            x_i = vit_block(P, sd, bi, x_i, cfg.heads, cfg.enc_eps, drops)
            x_a = vit_block(P, sd, ba, x_a, cfg.heads, cfg.enc_eps, drops)
        else:                                                                              # :104-107
            _, n_i = vit_block(P, sd, bi, torch.cat((x_f, x_i), 1), cfg.heads, cfg.enc_eps, drops).split((nF, nI), 1)
            _, n_a = vit_block(P, sd, ba, torch.cat((x_f, x_a), 1), cfg.heads, cfg.enc_eps, drops).split((nF, nA), 1)
            x_f = fusion_block(P, sd, f"encoder.fusion_blocks.{l}", x_f, x_i, x_a, cfg, drops)   # pre-block x_i / x_a
            x_i, x_a = n_i, n_a
        if return_embs:
            embs.append((x_i, x_a, x_f))
    x_i = P.layer_norm(x_i, sd["encoder.image.norm.weight"], sd["encoder.image.norm.bias"], cfg.enc_eps)     # :111
    x_a = P.layer_norm(x_a, sd["encoder.audio.norm.weight"], sd["encoder.audio.norm.bias"], cfg.enc_eps)     # :112
    x_f = P.layer_norm(x_f, sd["encoder.fusion_norm.weight"], sd["encoder.fusion_norm.bias"], cfg.fusion_eps)  # :113
    return (x_i, x_a, x_f, embs) if return_embs else (x_i, x_a, x_f)


# --------------------------------------------------------------------------------------
# a6: decoder  (avmae.py:147-180)
# --------------------------------------------------------------------------------------
def decoder_forward(P: _Prec, sd, cfg: OracleConfig, x: Tensor, x_fusion: Tensor, ids_restore: Tensor, modality: str) -> Tensor:
    m = modality
    B, nFus, nMask = x.shape[0], x_fusion.shape[1], ids_restore.shape[1] - x.shape[1]
    ew, eb = sd[f"{m}_decoder_embed.weight"], sd[f"{m}_decoder_embed.bias"]
    x, x_fusion = P.linear(x, ew, eb), P.linear(x_fusion, ew, eb)                         # :158
    mt = sd[f"{m}_decoder_mask_token"]
    x = torch.cat([x, mt.to(x.dtype).expand(B, nMask, -1)], 1)                            # :161
    x = x.gather(1, ids_restore.unsqueeze(-1).expand(-1, -1, x.shape[2]))                 # :162
    x = x + sd[f"{m}_decoder_pos_embed"]                                                  # :165
    x = torch.cat([x_fusion.to(x.dtype), x], 1)                                           # :169
    for i in range(cfg.dec_depth):                                                        # :170-171
        x = vit_block(P, sd, f"{m}_decoder_blocks.{i}", x, cfg.dec_heads, cfg.dec_eps)
    x = x[:, nFus:, :]                                                                    # :172
    x = P.layer_norm(x, sd[f"{m}_decoder_norm.weight"], sd[f"{m}_decoder_norm.bias"], cfg.dec_eps)
    return P.linear(x, sd[f"{m}_decoder_pred.weight"], sd[f"{m}_decoder_pred.bias"])     # :179


# --------------------------------------------------------------------------------------
# a7: patchify + loss  (avmae.py:183-214)
# --------------------------------------------------------------------------------------
def patchify(x: Tensor, patch: int) -> Tensor:
    """avmae.py:201-214: (N,C,H,W) -> (N, gH*gW, p*p*C), in-patch order (p, q, c)."""
    bs, c, H, W = x.shape
    gH, gW = H // patch, W // patch
    x = x.reshape(bs, c, gH, patch, gW, patch)
    x = torch.einsum("nchpwq->nhwpqc", x)
    return x.reshape(bs, gH * gW, patch * patch * c)


def forward_loss(target: Tensor, pred: Tensor, mask: Tensor, norm_pix_loss: bool) -> Tensor:
    """avmae.py:183-198.  Unbiased variance (:191), +1e-6 inside the sqrt."""
    if norm_pix_loss:
        mean = target.mean(dim=-1, keepdim=True)
        var = target.var(dim=-1, keepdim=True)
        target = (target - mean) / (var + 1.0e-6) ** 0.5
    loss = (pred - target) ** 2
    loss = loss.mean(dim=-1)
    return (loss * mask).sum() / mask.sum()


# --------------------------------------------------------------------------------------
# a8: AVMAE.forward  (avmae.py:216-236)
# --------------------------------------------------------------------------------------
def avmae_forward(sd: Dict[str, Tensor], cfg: OracleConfig, image: Tensor, audio: Tensor,
                  noise_image: Tensor, noise_audio: Tensor, amp: bool = False, sdpa: bool = False) -> Dict[str, Tensor]:
    """Returns a dict with the reference's 4 outputs plus the intermediate tensors parity tests
    compare (ids, masks, encoder outputs)."""
    P = _Prec(amp, sdpa)
    ik, im, ir = random_masking(noise_image, cfg.image_mask_ratio)                        # :220
    ak, am, ar = random_masking(noise_audio, cfg.audio_mask_ratio)                        # :221
    x_i, x_a, x_f = encoder_forward(P, sd, cfg, image, audio, ik, ak)                     # :224
    pred_i = decoder_forward(P, sd, cfg, x_i, x_f, ir, "image")                           # :228
    loss_i = forward_loss(patchify(image, cfg.patch), pred_i, im, cfg.image_norm_loss)    # :227,229
    pred_a = decoder_forward(P, sd, cfg, x_a, x_f, ar, "audio")                           # :233
    loss_a = forward_loss(patchify(audio, cfg.patch), pred_a, am, cfg.audio_norm_loss)    # :232,234
    return dict(loss_image=loss_i, loss_audio=loss_a, pred_image=pred_i, pred_audio=pred_a,
                x_image=x_i, x_audio=x_a, x_fusion=x_f,
                image_ids_keep=ik, image_mask=im, image_ids_restore=ir,
                audio_ids_keep=ak, audio_mask=am, audio_ids_restore=ar)


def loss_and_grads(sd: Dict[str, Tensor], cfg: OracleConfig, image, audio, noise_image, noise_audio, amp=False):
    """fwd + bwd of loss_image + loss_audio (train.py:164-165, misc.py:71-74 without a scaler).
    Returns (outputs, grads) with grads keyed like the state dict (frozen pos-embeds excluded)."""
    leaves = {k: (v.detach().clone().requires_grad_(k not in FROZEN_KEYS)) for k, v in sd.items()}
    out = avmae_forward(leaves, cfg, image, audio, noise_image, noise_audio, amp=amp)
    (out["loss_image"] + out["loss_audio"]).backward()
    grads = {k: v.grad for k, v in leaves.items() if v.requires_grad}
    return {k: (v.detach() if isinstance(v, Tensor) else v) for k, v in out.items()}, grads


# --------------------------------------------------------------------------------------
# a11: AVClassifier  (classifier.py:42-59)
# --------------------------------------------------------------------------------------
def classifier_state(cfg: OracleConfig, num_classes: int, seed: int = 0, input_norm: bool = False) -> Dict[str, Tensor]:
    """Encoder part of ``build_state`` (keys ``encoder.*``) plus the heads (classifier.py:20-22) and, with
    ``input_norm``, the BatchNorm1d buffers (:15-18)."""
    sd = {k: v for k, v in build_state(cfg, seed=seed).items() if k.startswith("encoder.")}
    g = torch.Generator().manual_seed(seed + 101)
    for m in ("image", "audio", "fusion"):
        sd[f"{m}_head.weight"] = torch.randn(num_classes, cfg.dim, generator=g) * 0.05
        sd[f"{m}_head.bias"] = torch.randn(num_classes, generator=g) * 0.05
        if input_norm:
            sd[f"{m}_norm.running_mean"] = torch.zeros(cfg.dim)
            sd[f"{m}_norm.running_var"] = torch.ones(cfg.dim)
            sd[f"{m}_norm.num_batches_tracked"] = torch.tensor(0)
    return sd


def classifier_forward(sd: Dict[str, Tensor], cfg: OracleConfig, image: Tensor, audio: Tensor, input_norm: bool = False,
                       training: bool = True, freeze_encoder: bool = False, momentum: float = 0.1, eps: float = 1e-6, drops=None):
    """Returns (pred_image, pred_audio, pred_fusion) and the updated BatchNorm running statistics.  The reference
    runs this path in fp32 (configs/linprobe.yaml:35, finetune.yaml:50)."""
    P = _Prec(False)
    if freeze_encoder:                                                                    # :43-45
        with torch.no_grad():
            xs = encoder_forward(P, sd, cfg, image, audio)
    else:
        xs = encoder_forward(P, sd, cfg, image, audio, drops=None if drops is None else iter(drops))   # :47
    preds, stats = [], {}
    for m, x in zip(("image", "audio", "fusion"), xs):
        f = x.mean(dim=1)                                                                 # :49
        if input_norm:                                                                    # :50-54, nn.BatchNorm1d(affine=False)
            rm, rv = sd[f"{m}_norm.running_mean"], sd[f"{m}_norm.running_var"]
            if training:
                mu, var = f.mean(0), f.var(0, unbiased=False)
                stats[f"{m}_norm.running_mean"] = (1 - momentum) * rm + momentum * mu.detach()
                stats[f"{m}_norm.running_var"] = (1 - momentum) * rv + momentum * f.var(0, unbiased=True).detach()
            else:
                mu, var = rm, rv
            f = (f - mu) / torch.sqrt(var + eps)
        preds.append(F.linear(f, sd[f"{m}_head.weight"], sd[f"{m}_head.bias"]))           # :56-58
    return tuple(preds), stats


def classifier_loss_and_grads(sd, cfg, image, audio, target_w, input_norm=False, training=True, freeze_encoder=False, drops=None):
    """Scalar = sum_m (pred_m * target_w).sum() -- a fixed linear functional of the three predictions, so that the
    gradient check does not depend on a loss the path does not own.  Returns (preds, stats, grads)."""
    trainable = lambda k: (k not in FROZEN_KEYS) and ("running_" not in k) and ("num_batches" not in k) and \
        not (freeze_encoder and k.startswith("encoder."))
    leaves = {k: (v.detach().clone().requires_grad_(True) if (v.is_floating_point() and trainable(k)) else v) for k, v in sd.items()}
    preds, stats = classifier_forward(leaves, cfg, image, audio, input_norm, training, freeze_encoder, drops=drops)
    sum((p * target_w).sum() for p in preds).backward()
    grads = {k: v.grad for k, v in leaves.items() if isinstance(v, Tensor) and v.requires_grad and v.grad is not None}
    return tuple(p.detach() for p in preds), stats, grads


# --------------------------------------------------------------------------------------
# a9: AdamW restatement (torch.optim.AdamW single-tensor math, used by train.py:93 with betas (.9,.95))
# --------------------------------------------------------------------------------------
def adamw_step(p: Tensor, g: Tensor, m: Tensor, v: Tensor, step: int, lr: float, beta1: float, beta2: float,
               eps: float, weight_decay: float):
    """In-place torch.optim.AdamW update (decoupled weight decay; bias-corrected;
    denom = sqrt(v)/sqrt(1-beta2^t) + eps).  ``step`` is 1-based."""
    p.mul_(1 - lr * weight_decay)
    m.mul_(beta1).add_(g, alpha=1 - beta1)
    v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    p.addcdiv_(m, denom, value=-lr / bc1)

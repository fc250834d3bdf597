"""Algorithmic FLOPs of the hot path per clip-pair (the figure ``roofline.achieved`` and
``step_tflops`` in bench.py are computed from).

Convention: 2*M*K*N per Linear / conv-as-GEMM, 2*H*Nq*Nk*(d_qk + d_v) per attention; backward = 2x forward except
that the patch embedding has no input gradient, so ``F_step = 3 * F_fwd - F_patch_embed``; the fusion-token rows of
the modality blocks (deepavfusion.py:104-105 discards them) are counted as keys / values only; every other GEMM is
counted as the reference writes it -- except, with ``factorised=True``, the pair attention's k / v projections
(fusion_blocks.py:245-258), which this build evaluates in the exactly-equivalent factorised form:
na + nv rows instead of na * nv, as the survey requires to be subtracted before quoting achieved FLOP/s.
No credit for recomputation.  tests/test_host_cpu.py pins the four BASELINE configurations to the survey's constants.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Tuple


@dataclass
class PathShape:
    dim: int = 768
    depth: int = 12
    heads: int = 12
    mlp_ratio: float = 4.0
    image_patches: int = 196          # 224 x 224 / 16^2
    audio_patches: int = 96           # 128 x 192 / 16^2
    image_chans: int = 3
    audio_chans: int = 1
    patch: int = 16
    fusion_tkns: Tuple[int, int, int] = (16, 8, 8)
    fusion_layers: int = 12
    fusion_attn_ratio: float = 0.25
    fusion_mlp_ratio: float = 1.0
    fusion_heads: int = 12
    dec_dim: int = 512
    dec_depth: int = 8
    dec_heads: int = 16
    dec_mlp_ratio: float = 4.0
    image_keep: int = 49              # int(196 * (1 - 0.75)); = image_patches when unmasked
    audio_keep: int = 19              # int(96 * (1 - 0.8))
    num_classes: int = 0              # > 0: classifier heads (configs 4-5) instead of the MAE decoders


def _lin(rows: int, k: int, n: int) -> float:
    return 2.0 * rows * k * n


def _attn_flops(heads: int, nq: int, nk: int, dqk: int, dv: int) -> float:
    return 2.0 * heads * nq * nk * (dqk + dv)


def forward_breakdown(s: PathShape, factorised: bool = True, attention: bool = True) -> dict:
    """Forward FLOPs per clip-pair by component.  ``attention=False`` leaves the QK^T / PV contractions out (the
    Linear / conv GEMMs alone: what the tcgen05 GEMM kernel executes)."""
    D, H = s.dim, s.heads
    _attn = _attn_flops if attention else (lambda *a: 0.0)
    hd = D // H
    nF = sum(s.fusion_tkns)
    nmm, nv, na = s.fusion_tkns
    out = {}
    out["patch_embed"] = (_lin(s.image_patches, s.image_chans * s.patch ** 2, D) + _lin(s.audio_patches, s.audio_chans * s.patch ** 2, D))

    def vit_block(n_live: int, n_prefix: int, dim: int, heads: int, mlp: float) -> float:
        S = n_live + n_prefix
        hidden = int(dim * mlp)
        return (_lin(n_live, dim, dim) + _lin(S, dim, 2 * dim) + _attn(heads, n_live, S, dim // heads, dim // heads)
                + _lin(n_live, dim, dim) + _lin(n_live, dim, hidden) + _lin(n_live, hidden, dim))
    n_fused = min(s.fusion_layers, s.depth)
    out["image_blocks"] = n_fused * vit_block(s.image_keep, nF, D, H, s.mlp_ratio) + (s.depth - n_fused) * vit_block(s.image_keep, 0, D, H, s.mlp_ratio)
    out["audio_blocks"] = n_fused * vit_block(s.audio_keep, nF, D, H, s.mlp_ratio) + (s.depth - n_fused) * vit_block(s.audio_keep, 0, D, H, s.mlp_ratio)

    qk = int(D * s.fusion_attn_ratio)
    fh = s.fusion_heads
    cross = lambda nq, nk: _lin(nq, D, D) + _lin(nk, D, 2 * D) + _attn(fh, nq, nk, D // fh, D // fh) + _lin(nq, D, D)
    pair_rows = (nv + na) if factorised else nv * na
    pair_k = D if factorised else 2 * D           # factorised: k(v_i) + k(a_j), each a D-wide Linear
    pair = (_lin(nmm, D, qk) + _lin(pair_rows, pair_k, qk) + _lin(pair_rows, pair_k, D)
            + _attn(fh, nmm, nv * na, qk // fh, D // fh) + _lin(nmm, D, D))
    fmlp = _lin(nF, D, int(D * s.fusion_mlp_ratio)) + _lin(nF, int(D * s.fusion_mlp_ratio), D)
    out["fusion_blocks"] = n_fused * (cross(nv, s.image_keep) + cross(na, s.audio_keep) + pair + fmlp)

    if s.num_classes > 0:
        out["heads"] = 3 * _lin(1, D, s.num_classes)
        return out
    Dd = s.dec_dim
    out["decoder_embed"] = _lin(s.image_keep + nF, D, Dd) + _lin(s.audio_keep + nF, D, Dd)
    out["image_decoder"] = s.dec_depth * vit_block(s.image_patches + nF, 0, Dd, s.dec_heads, s.dec_mlp_ratio)
    out["audio_decoder"] = s.dec_depth * vit_block(s.audio_patches + nF, 0, Dd, s.dec_heads, s.dec_mlp_ratio)
    out["pred"] = _lin(s.image_patches, Dd, s.image_chans * s.patch ** 2) + _lin(s.audio_patches, Dd, s.audio_chans * s.patch ** 2)
    return out


def gflop_forward(s: PathShape, factorised: bool = True, attention: bool = True) -> float:
    return sum(forward_breakdown(s, factorised, attention).values()) / 1e9


def gflop_step(s: PathShape, factorised: bool = True, attention: bool = True) -> float:
    """fwd + bwd of one clip-pair: 3 * F_fwd - F_patch_embed (no dX through the patch embedding)."""
    b = forward_breakdown(s, factorised, attention)
    return (3.0 * sum(b.values()) - b["patch_embed"]) / 1e9


# BASELINE.json configs[1..4]
def vggsound_pretrain() -> PathShape:
    return PathShape(fusion_attn_ratio=0.25, fusion_mlp_ratio=1.0)


def audioset_pretrain() -> PathShape:
    return PathShape(fusion_attn_ratio=1.0, fusion_mlp_ratio=4.0)


def unmasked_classifier(num_classes: int = 310, fusion_attn_ratio: float = 0.25, fusion_mlp_ratio: float = 1.0) -> PathShape:
    return PathShape(fusion_attn_ratio=fusion_attn_ratio, fusion_mlp_ratio=fusion_mlp_ratio, image_keep=196, audio_keep=96,
                     num_classes=num_classes)

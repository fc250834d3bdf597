"""Shared helpers for the model-level parity tests (CPU-emulated and GPU)."""
from __future__ import annotations

import contextlib
from functools import partial

import torch
from torch import nn

from oracle import avmae_oracle as O


def tiny_cfg(**kw) -> O.OracleConfig:
    base = dict(image_size=(64, 64), audio_size=(32, 96), dim=128, depth=2, heads=2, fusion_heads=2,
                dec_dim=128, dec_depth=2, dec_heads=4, fusion_attn_ratio=0.25, fusion_mlp_ratio=1.0)
    base.update(kw)
    return O.OracleConfig(**base)


def build_model(cfg: O.OracleConfig, device="cpu", drop_path=0.0):
    """Our drop-in AVMAE(DeepAVFusion) for an oracle config (ViT-B or a tiny test size)."""
    from deepavfusion_b200.models import AVMAE, DeepAVFusion, vits

    def arch(**kw):
        kw.pop("pretrained", None)
        return vits.ViT(patch_size=cfg.patch, embed_dim=cfg.dim, depth=cfg.depth, num_heads=cfg.heads, mlp_ratio=cfg.mlp_ratio,
                        norm_layer=partial(nn.LayerNorm, eps=cfg.enc_eps), **kw)
    vits.__dict__["vit_test"] = arch
    enc = DeepAVFusion(image_arch="vit_test", image_pretrained="", image_size=cfg.image_size,
                       audio_arch="vit_test", audio_pretrained="", audio_size=cfg.audio_size,
                       fusion_layers=cfg.fusion_layers, num_fusion_tkns=cfg.fusion_tkns,
                       fusion_mlp_ratio=cfg.fusion_mlp_ratio, fusion_attn_ratio=cfg.fusion_attn_ratio, fusion_num_heads=cfg.fusion_heads,
                       drop_path=drop_path)
    model = AVMAE(enc, enc.embed_dim, image_decoder_depth=cfg.dec_depth, image_mask_ratio=cfg.image_mask_ratio,
                  image_norm_loss=cfg.image_norm_loss, audio_decoder_depth=cfg.dec_depth, audio_mask_ratio=cfg.audio_mask_ratio,
                  audio_norm_loss=cfg.audio_norm_loss, decoder_dim=cfg.dec_dim, num_heads=cfg.dec_heads, mlp_ratio=cfg.dec_mlp_ratio)
    return model.to(device)


def build_classifier(cfg: O.OracleConfig, num_classes: int, freeze_encoder: bool, input_norm: bool, device="cpu", drop_path=0.0):
    """Our drop-in AVClassifier(DeepAVFusion) for an oracle config."""
    from deepavfusion_b200.models import AVClassifier
    enc = build_model(cfg, "cpu", drop_path=drop_path).encoder
    return AVClassifier(enc, num_classes, freeze_encoder=freeze_encoder, input_norm=input_norm).to(device)


@contextlib.contextmanager
def inject_droppath(scales):
    """Make the model's DropPath draws (functional.droppath_scale) return the given per-sample scale vectors in order."""
    import deepavfusion_b200.functional as Fn
    it, orig = iter(scales), Fn.droppath_scale

    def fake(B, drop_prob, device):
        s = next(it)
        assert s.numel() == B
        return s.to(device=device, dtype=torch.float32)
    Fn.droppath_scale = fake
    try:
        yield it
    finally:
        Fn.droppath_scale = orig



def make_inputs(cfg, B, seed=1):
    g = torch.Generator().manual_seed(seed)
    image = torch.randn(B, cfg.image_chans, *cfg.image_size, generator=g)
    audio = torch.randn(B, cfg.audio_chans, *cfg.audio_size, generator=g)
    return image, audio


def make_noise(cfg, B, seed=2):
    g = torch.Generator().manual_seed(seed)
    return (torch.rand(B, cfg.image_grid[0] * cfg.image_grid[1], generator=g),
            torch.rand(B, cfg.audio_grid[0] * cfg.audio_grid[1], generator=g))


@contextlib.contextmanager
def inject_rand(noises):
    """Make the two torch.rand(N, L, device=...) draws of AVMAE.random_masking return our noise."""
    noises = list(noises)
    orig = torch.rand

    def fake(*size, **kw):
        n = noises.pop(0)
        assert tuple(size) == tuple(n.shape), (size, n.shape)
        return n.clone().to(kw.get("device", "cpu"))
    torch.rand = fake
    try:
        yield
    finally:
        torch.rand = orig


def grad_report(named_params, ref_grads, amp_grads=None):
    """Per-tensor rel-L2 error of .grad against the fp32 oracle.  Returns (failures, worst, global rel)."""
    gn = torch.sqrt(sum((g.double() ** 2).sum() for g in ref_grads.values())).item()
    atol = 1e-5 * gn                                  # zero-gradient K-biases (SURVEY.md 7.1 corollary)
    failures, worst = [], 0.0
    num = den = 0.0
    for k, ref in ref_grads.items():
        g = named_params[k].grad
        assert g is not None, f"no gradient for {k}"
        g = g.detach().float().cpu()
        diff = (g - ref).norm().item()
        num += diff ** 2
        den += ref.norm().item() ** 2
        if ref.norm().item() < atol * 10:
            ok = diff <= atol * 10
            rel = diff
        else:
            rel = diff / ref.norm().item()
            gate = 3e-2
            if amp_grads is not None:
                gate = max(gate, 1.5 * ((amp_grads[k].float() - ref).norm() / ref.norm()).item())
            ok = rel <= gate
            worst = max(worst, rel)
        if not ok:
            failures.append((k, rel))
    return failures, worst, (num / max(den, 1e-30)) ** 0.5

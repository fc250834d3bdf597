"""Generate tests/golden/* from the REAL reference (run in the build container only).

TEST INFRASTRUCTURE ONLY.  Imports /root/reference/models/{deepavfusion,avmae,...}.py through
oracle/timm_shim, checks that oracle/avmae_oracle.py reproduces the reference's outputs and
gradients on identical weights / inputs / mask noise, and stores small fixtures that travel to
the GPU box (where /root/reference does not exist):

    python oracle/make_golden.py            # writes tests/golden/*.json, *.npz

Fixture recipe (all CPU fp32): weights = oracle.build_state(cfg, seed=0) loaded into the
reference model with load_state_dict(strict=True); inputs = Generator(1) randn; mask noise =
Generator(2) rand, injected into the reference by patching torch.rand for the two calls at
avmae.py:127.
"""
from __future__ import annotations

import hashlib
import json
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("DAVF_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(HERE, "timm_shim"))
sys.path.insert(0, REF)

from oracle import avmae_oracle as O  # noqa: E402


def tensor_digest(t: torch.Tensor) -> str:
    return hashlib.sha256(t.detach().contiguous().cpu().numpy().tobytes()).hexdigest()[:16]


def build_reference(cfg: O.OracleConfig):
    from models.deepavfusion import DeepAVFusion   # reference
    from models.avmae import AVMAE                 # reference
    enc = DeepAVFusion(
        image_arch="vit_base", image_pretrained="", image_size=cfg.image_size,
        audio_arch="vit_base", audio_pretrained="", audio_size=cfg.audio_size,
        fusion_arch="factorized_mmi", fusion_layers=cfg.fusion_layers, num_fusion_tkns=cfg.fusion_tkns,
        fusion_mlp_ratio=cfg.fusion_mlp_ratio, fusion_attn_ratio=cfg.fusion_attn_ratio, fusion_num_heads=cfg.fusion_heads)
    model = AVMAE(enc, enc.embed_dim,
                  image_decoder_arch="plain", image_decoder_depth=cfg.dec_depth, image_mask_ratio=cfg.image_mask_ratio,
                  image_norm_loss=cfg.image_norm_loss,
                  audio_decoder_arch="plain", audio_decoder_depth=cfg.dec_depth, audio_mask_ratio=cfg.audio_mask_ratio,
                  audio_norm_loss=cfg.audio_norm_loss)
    return model


class _InjectRand:
    """Make the reference's two torch.rand(N, L) calls (avmae.py:127) return our noise."""

    def __init__(self, noises):
        self.noises, self.i = list(noises), 0

    def __enter__(self):
        self._orig = torch.rand

        def fake(*size, **kw):
            n = self.noises[self.i]
            self.i += 1
            assert tuple(size) == tuple(n.shape), (size, n.shape)
            return n.clone()
        torch.rand = fake
        return self

    def __exit__(self, *a):
        torch.rand = self._orig


def make_inputs(cfg, B, seed=1):
    g = torch.Generator().manual_seed(seed)
    image = torch.randn(B, cfg.image_chans, *cfg.image_size, generator=g)
    audio = torch.randn(B, cfg.audio_chans, *cfg.audio_size, generator=g)
    return image, audio


def make_noise(cfg, B, seed=2):
    g = torch.Generator().manual_seed(seed)
    Li = cfg.image_grid[0] * cfg.image_grid[1]
    La = cfg.audio_grid[0] * cfg.audio_grid[1]
    return torch.rand(B, Li, generator=g), torch.rand(B, La, generator=g)


def rel(a, b):
    return float((a - b).norm() / (b.norm() + 1e-30))


def run_config(name, cfg, B, out_dir):
    torch.manual_seed(0)
    t0 = time.time()
    model = build_reference(cfg)
    ref_keys = {k: list(v.shape) for k, v in model.state_dict().items()}
    n_params = sum(p.numel() for p in model.parameters())
    shapes = O.state_shapes(cfg)
    assert set(shapes) == set(ref_keys), (set(shapes) ^ set(ref_keys))
    for k in shapes:
        assert list(shapes[k]) == ref_keys[k], (k, shapes[k], ref_keys[k])
    frozen = sorted(k for k, p in model.named_parameters() if not p.requires_grad)
    assert tuple(frozen) == tuple(sorted(O.FROZEN_KEYS)), frozen

    # the reference's own sin-cos pos-embeds must equal the oracle's restatement
    sd0 = model.state_dict()
    for k in shapes:
        if k.endswith("pos_embed"):
            grid = cfg.image_grid if "image" in k else cfg.audio_grid
            mine = torch.from_numpy(O.sincos_2d(shapes[k][-1], grid)).float().unsqueeze(0)
            assert torch.equal(mine, sd0[k]), k

    sd = O.build_state(cfg, seed=0)
    model.load_state_dict(sd, strict=True)
    image, audio = make_inputs(cfg, B)
    ni, na = make_noise(cfg, B)

    model.train()
    with _InjectRand([ni, na]):
        li, la, pi, pa = model(image, audio)
    (li + la).backward()
    ref_grads = {k: p.grad for k, p in model.named_parameters() if p.requires_grad}
    assert all(g is not None for g in ref_grads.values())

    out, grads = O.loss_and_grads(sd, cfg, image, audio, ni, na)
    # --- oracle == reference ---
    assert abs(float(out["loss_image"]) - float(li)) <= 1e-6 * abs(float(li)), (out["loss_image"], li)
    assert abs(float(out["loss_audio"]) - float(la)) <= 1e-6 * abs(float(la)), (out["loss_audio"], la)
    assert rel(out["pred_image"], pi) < 1e-5 and rel(out["pred_audio"], pa) < 1e-5
    worst = max(rel(grads[k], ref_grads[k]) if ref_grads[k].norm() > 1e-7 else float((grads[k] - ref_grads[k]).abs().max())
                for k in ref_grads)
    assert set(grads) == set(ref_grads)
    assert worst < 2e-4, worst
    print(f"[{name}] oracle == reference: loss {float(li):.6f}/{float(la):.6f}, worst grad rel err {worst:.2e}")

    # encoder-only unmasked forward (AVMAE.forward_encoder, avmae.py:144-145)
    with torch.no_grad():
        rxi, rxa, rxf = model.forward_encoder(image, audio)
        oxi, oxa, oxf = O.encoder_forward(O._Prec(False), sd, cfg, image, audio)
    assert rel(oxi, rxi) < 1e-5 and rel(oxa, rxa) < 1e-5 and rel(oxf, rxf) < 1e-5

    # mask pieces straight from the reference's random_masking
    with _InjectRand([ni]):
        r_keep, r_mask, r_restore = model.random_masking(B, ni.shape[1], cfg.image_mask_ratio, device="cpu")
    assert torch.equal(r_keep, out["image_ids_keep"]) and torch.equal(r_restore, out["image_ids_restore"])
    assert torch.equal(r_mask, out["image_mask"])

    gn = torch.sqrt(sum((g.double() ** 2).sum() for g in ref_grads.values()))
    keys = sorted(ref_grads)
    np.savez_compressed(
        os.path.join(out_dir, f"{name}.npz"),
        noise_image=ni.numpy(), noise_audio=na.numpy(),
        image_ids_keep=r_keep.numpy(), image_ids_restore=r_restore.numpy(), image_mask=r_mask.numpy(),
        audio_ids_keep=out["audio_ids_keep"].numpy(), audio_ids_restore=out["audio_ids_restore"].numpy(),
        audio_mask=out["audio_mask"].numpy(),
        loss_image=np.float64(float(li)), loss_audio=np.float64(float(la)),
        pred_image_head=pi.detach()[:, :4, :16].numpy(), pred_audio_head=pa.detach()[:, :4, :16].numpy(),
        enc_x_fusion_unmasked=rxf.numpy().astype(np.float32),
        enc_x_image_unmasked_head=rxi[:, :4, :32].numpy(), enc_x_audio_unmasked_head=rxa[:, :4, :32].numpy(),
        grad_norm_global=np.float64(float(gn)),
        grad_norms=np.array([float(ref_grads[k].double().norm()) for k in keys]),
        grad_heads=np.stack([np.pad(ref_grads[k].flatten()[:8].numpy(), (0, max(0, 8 - ref_grads[k].numel()))) for k in keys]),
    )
    meta = dict(
        name=name, batch=B, n_params=int(n_params), n_state_keys=len(ref_keys), grad_keys=keys,
        state_shapes=ref_keys, frozen=frozen,
        cfg=dict(fusion_attn_ratio=cfg.fusion_attn_ratio, fusion_mlp_ratio=cfg.fusion_mlp_ratio,
                 fusion_tkns=list(cfg.fusion_tkns), fusion_layers=cfg.fusion_layers),
        digests=dict(image=tensor_digest(image), audio=tensor_digest(audio),
                     state=tensor_digest(torch.cat([sd[k].flatten()[:64] for k in sorted(sd)]))),
        recipe="weights oracle.build_state(seed=0); inputs Generator(1) randn; noise Generator(2) rand; CPU fp32",
        torch=torch.__version__, seconds=round(time.time() - t0, 1),
    )
    with open(os.path.join(out_dir, f"{name}.json"), "w") as f:
        json.dump(meta, f, indent=1)
    return meta


def classifier_fixture(out_dir):
    """a11: the oracle's AVClassifier restatement against the REAL reference classifier (models/classifier.py) on a
    small encoder (dims of tests/model_utils.tiny_cfg), lin-probe (frozen, BatchNorm) and fine-tune (trainable)."""
    from functools import partial
    from models.deepavfusion import DeepAVFusion   # reference
    from models.classifier import AVClassifier     # reference
    import models.vits as ref_vits                 # reference
    cfg = O.OracleConfig(image_size=(64, 64), audio_size=(32, 96), dim=128, depth=2, heads=2, fusion_heads=2,
                         dec_dim=128, dec_depth=2, dec_heads=4, fusion_attn_ratio=0.25, fusion_mlp_ratio=1.0)   # = tests/model_utils.tiny_cfg()

    def ref_arch(pretrained=None, **kw):           # a small reference ViT (the reference only registers base / large / huge)
        return ref_vits.ViT(patch_size=cfg.patch, embed_dim=cfg.dim, depth=cfg.depth, num_heads=cfg.heads, mlp_ratio=cfg.mlp_ratio,
                            norm_layer=partial(torch.nn.LayerNorm, eps=cfg.enc_eps), **kw)
    ref_vits.__dict__["vit_test"] = ref_arch
    C, B = 10, 4
    image, audio = make_inputs(cfg, B)
    g = torch.Generator().manual_seed(5)
    tw = torch.randn(B, C, generator=g)
    saved = {}
    import timm.models.layers as shim_layers      # oracle/timm_shim: DropPath as the reference imports it
    n_draws = 6 * cfg.depth                         # per layer: image block 2, audio block 2, fusion block 2 (reference call order)
    gd = torch.Generator().manual_seed(9)
    drops = [(torch.rand(B, generator=gd) < 0.8).float() / 0.8 for _ in range(n_draws)]
    saved["droppath_scales"] = torch.stack(drops).numpy()
    for tag, freeze, inorm in (("linprobe", True, True), ("finetune", False, False), ("finetune_droppath", False, False)):
        dp = 0.2 if tag == "finetune_droppath" else 0.0
        torch.manual_seed(0)
        enc = DeepAVFusion(image_arch="vit_test", image_pretrained="", image_size=cfg.image_size,
                           audio_arch="vit_test", audio_pretrained="", audio_size=cfg.audio_size,
                           fusion_arch="factorized_mmi", fusion_layers=cfg.fusion_layers, num_fusion_tkns=cfg.fusion_tkns,
                           fusion_mlp_ratio=cfg.fusion_mlp_ratio, fusion_attn_ratio=cfg.fusion_attn_ratio, fusion_num_heads=cfg.fusion_heads,
                           drop_path=dp)
        model = AVClassifier(enc, C, freeze_encoder=freeze, input_norm=inorm)
        sd = O.classifier_state(cfg, C, seed=0, input_norm=inorm)
        assert set(sd) == set(model.state_dict()), set(sd) ^ set(model.state_dict())
        model.load_state_dict(sd, strict=True)
        model.train()
        it = iter(drops)
        orig_fwd = shim_layers.DropPath.forward

        def injected(self, x):                      # the reference's DropPath modules, fed our masks in call order
            if self.drop_prob == 0. or not self.training:
                return x
            return x * next(it).view(-1, *([1] * (x.ndim - 1)))
        shim_layers.DropPath.forward = injected
        try:
            preds = model(image, audio)
        finally:
            shim_layers.DropPath.forward = orig_fwd
        if dp:
            assert next(it, None) is None, "the reference drew a different number of DropPath masks"
        sum((p * tw).sum() for p in preds).backward()
        ref_grads = {k: p.grad for k, p in model.named_parameters() if p.requires_grad and p.grad is not None}
        opreds, stats, grads = O.classifier_loss_and_grads(sd, cfg, image, audio, tw, input_norm=inorm, training=True, freeze_encoder=freeze,
                                                           drops=drops if dp else None)
        for a, b in zip(opreds, preds):
            assert rel(a, b.detach()) < 1e-5, (tag, rel(a, b.detach()))
        assert set(grads) == set(ref_grads), (tag, set(grads) ^ set(ref_grads))
        worst = max(rel(grads[k], ref_grads[k]) if ref_grads[k].norm() > 1e-7 else float((grads[k] - ref_grads[k]).abs().max()) for k in ref_grads)
        assert worst < 2e-4, (tag, worst)
        msd = model.state_dict()
        for k, v in stats.items():
            assert rel(v, msd[k]) < 1e-5, (tag, k)
        print(f"[classifier/{tag}] oracle == reference: preds ok, {len(grads)} grads, worst rel err {worst:.2e}")
        for n, pr in zip(("image", "audio", "fusion"), preds):
            saved[f"{tag}_pred_{n}"] = pr.detach().numpy()
        keys = sorted(ref_grads)
        saved[f"{tag}_grad_norms"] = np.array([float(ref_grads[k].double().norm()) for k in keys])
        saved[f"{tag}_grad_keys"] = np.array(keys)
        for k, v in stats.items():
            saved[f"{tag}_{k}"] = msd[k].numpy()
    saved["target_w"] = tw.numpy()
    np.savez_compressed(os.path.join(out_dir, "classifier_tiny.npz"), **saved)


def mask_ties_fixture(out_dir):
    """Ties: the reference calls argsort without stable=True (avmae.py:128-129); on CPU its result
    is recorded here next to the stable definition the build uses."""
    model_cls = __import__("models.avmae", fromlist=["AVMAE"]).AVMAE
    g = torch.Generator().manual_seed(7)
    noise = (torch.rand(16, 196, generator=g) * 12).floor() / 12.0      # many exact ties
    with _InjectRand([noise]):
        keep, mask, restore = model_cls.random_masking(None, 16, 196, 0.75, device="cpu")
    okeep, omask, orestore = O.random_masking(noise, 0.75)
    same = bool(torch.equal(keep, okeep) and torch.equal(restore, orestore) and torch.equal(mask, omask))
    np.savez_compressed(os.path.join(out_dir, "mask_ties.npz"), noise=noise.numpy(),
                        ids_keep=okeep.numpy(), ids_restore=orestore.numpy(), mask=omask.numpy(),
                        reference_cpu_agrees=np.bool_(same))
    print(f"[mask_ties] reference CPU argsort agrees with stable definition: {same}")


def main():
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    if len(sys.argv) > 1 and sys.argv[1] == "classifier":      # only the (cheap) classifier fixture
        classifier_fixture(out_dir)
        return
    torch.set_num_threads(len(os.sched_getaffinity(0)))
    m1 = run_config("vggsound_b2", O.OracleConfig(fusion_attn_ratio=0.25, fusion_mlp_ratio=1.0), 2, out_dir)
    assert m1["n_params"] == 320_563_712, m1["n_params"]             # SURVEY.md 8(c) KAT
    m2 = run_config("audioset_b1", O.OracleConfig(fusion_attn_ratio=1.0, fusion_mlp_ratio=4.0), 1, out_dir)
    assert m2["n_params"] == 378_997_760, m2["n_params"]
    m3 = run_config("sparse_fusion_b1", O.OracleConfig(fusion_attn_ratio=0.25, fusion_mlp_ratio=1.0, fusion_layers="0-3-7"), 1, out_dir)
    mask_ties_fixture(out_dir)
    classifier_fixture(out_dir)
    print("param counts:", m1["n_params"], m2["n_params"], m3["n_params"])


if __name__ == "__main__":
    main()

"""-m gpu: parity at PRODUCTION shapes.  The CPU oracle runs on the box's host cores (seconds at these sizes) and
EVERY gradient tensor of the CUDA path is compared with it element-wise (``model_utils.grad_report``: per-tensor
rel-L2 of the difference, gate max(3e-2, 1.5x the oracle's own bf16-autocast error), absolute gate for the
mathematically-zero K-bias gradients, global rel-L2 <= 2e-2; losses rtol 2e-3; predictions rel-L2 2e-2).

The tiny-dimension tests never reach the kernels the benchmark runs on: the CTA-pair GEMM
(``gemm_tc_kernel<256, ..., CG = 2>``) needs M, N >= 256 with >= 36 pair tiles, the tcgen05 attention needs
>= 8 query rows.  Each case below therefore also asserts, through ``davf_launch_count_kind``, that those kernel
families were the ones that served it.

  vggsound_b64   BASELINE configs[1]: r = 0.25 / mlp 1, 64 pairs (the bench workload)   avmae.py:216-236
  audioset_b16   configs[2] widths: r = 1 / mlp 4 (full-width fusion blocks)             fusion_blocks.py:216-289
  sparse_b16     fusion_layers = '0-3-7' (deepavfusion.py:38-46): blocks without fusion
  classifier     configs[3]/[4]: unmasked AVClassifier at ViT-B, fine-tune fwd+bwd       classifier.py:42-59
"""
import os

import pytest
import torch

from oracle import avmae_oracle as O
import model_utils as U

pytestmark = pytest.mark.gpu
REQUIRE_ATTN_TC = True


def _kinds():
    import deepavfusion_b200.kernels as K
    return {k: K.launch_count_kind(k) for k in (K.KIND_GEMM_2CTA, K.KIND_ATTN_TC, K.KIND_ATTN_MMA)}


@pytest.mark.parametrize("name,kw,B", [
    ("vggsound_b64", dict(fusion_attn_ratio=0.25, fusion_mlp_ratio=1.0), 64),
    ("audioset_b16", dict(fusion_attn_ratio=1.0, fusion_mlp_ratio=4.0), 16),
    ("sparse_b16", dict(fusion_attn_ratio=0.25, fusion_mlp_ratio=1.0, fusion_layers="0-3-7"), 16),
])
def test_vitb_every_gradient_elementwise(name, kw, B):
    import deepavfusion_b200.kernels as K
    torch.set_num_threads(len(os.sched_getaffinity(0)))
    cfg = O.OracleConfig(**kw)
    sd = O.build_state(cfg, seed=0)
    image, audio = U.make_inputs(cfg, B, seed=3)
    ni, na = U.make_noise(cfg, B, seed=4)
    out, grads = O.loss_and_grads(sd, cfg, image, audio, ni, na)
    _, amp_grads = O.loss_and_grads(sd, cfg, image, audio, ni, na, amp=True)

    model = U.build_model(cfg, "cuda")
    model.load_state_dict(sd, strict=True)
    before = _kinds()
    with U.inject_rand([ni, na]):
        li, la, pi, pa = model(image.cuda(), audio.cuda())
    (li + la).backward()
    model._davf_store.join_side_streams(torch.cuda.current_stream())
    torch.cuda.synchronize()
    after = _kinds()

    # masks: bit-exact
    with U.inject_rand([ni, na]):
        ik, im, ir = model.random_masking(B, ni.shape[1], cfg.image_mask_ratio, "cuda")
        ak, am, ar = model.random_masking(B, na.shape[1], cfg.audio_mask_ratio, "cuda")
    for got, ref in ((ik, "image_ids_keep"), (im, "image_mask"), (ir, "image_ids_restore"),
                     (ak, "audio_ids_keep"), (am, "audio_mask"), (ar, "audio_ids_restore")):
        assert torch.equal(got.cpu(), out[ref]), ref

    assert abs(li.item() - out["loss_image"].item()) <= 2e-3 * abs(out["loss_image"].item()), (li.item(), out["loss_image"].item())
    assert abs(la.item() - out["loss_audio"].item()) <= 2e-3 * abs(out["loss_audio"].item()), (la.item(), out["loss_audio"].item())
    rel = lambda a, b: ((a.float().cpu() - b).norm() / b.norm()).item()
    assert rel(pi, out["pred_image"]) < 2e-2 and rel(pa, out["pred_audio"]) < 2e-2
    named = dict(model.named_parameters())
    assert set(grads) == {k for k, p in named.items() if p.requires_grad}
    failures, worst, glob = U.grad_report(named, grads, amp_grads)
    assert not failures, f"{name}: {len(failures)} of {len(grads)} gradient tensors out of tolerance, e.g. {failures[:5]}"
    assert glob <= 2e-2, glob
    # the production kernel families served this case
    assert after[K.KIND_GEMM_2CTA] - before[K.KIND_GEMM_2CTA] >= 100, (before, after)
    assert not REQUIRE_ATTN_TC or after[K.KIND_ATTN_TC] - before[K.KIND_ATTN_TC] >= 40, (before, after)
    print(f"{name}: worst per-tensor rel-L2 {worst:.3e}, global {glob:.3e}; "
          f"2-CTA GEMM launches {after[K.KIND_GEMM_2CTA] - before[K.KIND_GEMM_2CTA]}, "
          f"tcgen05 attention launches {after[K.KIND_ATTN_TC] - before[K.KIND_ATTN_TC]}, "
          f"mma.sync attention launches {after[K.KIND_ATTN_MMA] - before[K.KIND_ATTN_MMA]}")


@pytest.mark.parametrize("tag,freeze,inorm,B", [("finetune_b32", False, False, 32), ("linprobe_b64", True, True, 64)])
def test_vitb_classifier_elementwise(tag, freeze, inorm, B):
    """BASELINE configs[3] / [4] at ViT-B: unmasked encoder (196 + 96 tokens, the long-sequence attention shapes) + the
    classifier tail, predictions and every gradient against the oracle."""
    import deepavfusion_b200.kernels as K
    torch.set_num_threads(len(os.sched_getaffinity(0)))
    C = 310
    cfg = O.OracleConfig(fusion_attn_ratio=0.25, fusion_mlp_ratio=1.0)
    sd = O.classifier_state(cfg, C, seed=0, input_norm=inorm)
    image, audio = U.make_inputs(cfg, B, seed=5)
    tw = torch.randn(B, C, generator=torch.Generator().manual_seed(6))
    preds, stats, grads = O.classifier_loss_and_grads(sd, cfg, image, audio, tw, input_norm=inorm, training=True, freeze_encoder=freeze)
    model = U.build_classifier(cfg, C, freeze, inorm, "cuda")
    model.load_state_dict(sd, strict=True)
    model.train()
    before = _kinds()
    out = model(image.cuda(), audio.cuda())
    sum((p * tw.cuda()).sum() for p in out).backward()
    torch.cuda.synchronize()
    after = _kinds()
    for p, r in zip(out, preds):
        assert float((p.detach().cpu() - r).norm() / r.norm()) < 2e-2
    named = dict(model.named_parameters())
    gnorm = float(torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values())))
    bad = []
    for k, r in grads.items():
        g = named[k].grad.detach().cpu()
        err = float((g - r).norm())
        if err > 6e-2 * float(r.norm()) + 1e-4 * gnorm:
            bad.append((k, err, float(r.norm())))
    assert not bad, f"{tag}: {len(bad)} of {len(grads)} gradient tensors out of tolerance, e.g. {bad[:5]}"
    num = sum(float((named[k].grad.detach().cpu() - r).norm()) ** 2 for k, r in grads.items())
    assert (num ** 0.5) / gnorm <= 2e-2
    assert after[K.KIND_GEMM_2CTA] - before[K.KIND_GEMM_2CTA] >= 50, (before, after)
    assert not REQUIRE_ATTN_TC or after[K.KIND_ATTN_TC] - before[K.KIND_ATTN_TC] >= 24, (before, after)

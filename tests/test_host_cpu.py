"""CPU: host-side orchestration (autograd wiring, gradient routing into the flat buffers, state_dict
contract, fail-loud behaviour) with the CUDA kernels replaced by their torch emulations, against the
oracle.  No CUDA compute is involved; the kernels themselves are checked by the -m gpu tests."""
import ctypes
import os
import re

import pytest
import torch

from oracle import avmae_oracle as O
import cpu_kernels
import model_utils as U

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture
def emulated(monkeypatch):
    cpu_kernels.install(monkeypatch)


@pytest.mark.parametrize("kw,B", [({}, 3), (dict(fusion_attn_ratio=1.0, fusion_mlp_ratio=4.0), 2),
                                  (dict(fusion_layers="1", depth=3), 2), (dict(fusion_layers="none"), 2)])
def test_model_orchestration_vs_oracle(emulated, kw, B):
    cfg = U.tiny_cfg(**kw)
    sd = O.build_state(cfg, seed=0)
    image, audio = U.make_inputs(cfg, B)
    ni, na = U.make_noise(cfg, B)
    out, grads = O.loss_and_grads(sd, cfg, image, audio, ni, na)
    _, amp_grads = O.loss_and_grads(sd, cfg, image, audio, ni, na, amp=True)
    model = U.build_model(cfg)
    model.load_state_dict(sd, strict=True)
    with U.inject_rand([ni, na]):
        li, la, pi, pa = model(image, audio)
    (li + la).backward()
    assert abs(li.item() - out["loss_image"].item()) < 2e-3 * out["loss_image"].item()
    assert abs(la.item() - out["loss_audio"].item()) < 2e-3 * out["loss_audio"].item()
    named = dict(model.named_parameters())
    if kw.get("fusion_layers") == "none":       # fusion tokens only feed the decoders then
        grads = {k: v for k, v in grads.items() if v is not None}
    failures, worst, glob = U.grad_report(named, grads, amp_grads)
    assert not failures, failures[:5]
    assert glob < 2e-2
    # gradients live in the flat buffer, frozen pos-embeds have none
    st = model._davf_store
    for k, p in named.items():
        if p.requires_grad:
            assert p.grad.data_ptr() == st.grad(p).data_ptr()
    assert named["encoder.image.pos_embed"].grad is None


@pytest.mark.parametrize("tag,freeze,inorm,dp", [("linprobe", True, True, 0.0), ("finetune", False, False, 0.0), ("finetune_bn", False, True, 0.0),
                                                 ("finetune_droppath", False, False, 0.2)])
def test_classifier_orchestration_vs_oracle(emulated, tag, freeze, inorm, dp):
    """a11 (classifier.py:42-59): drop-in AVClassifier, emulated kernels, against the oracle restatement; the
    fine-tuning config's stochastic depth (configs/finetune.yaml:36) with the DropPath masks injected on both sides."""
    _classifier_case("cpu", freeze, inorm, dp=dp)


def _classifier_case(device, freeze, inorm, C=10, B=4, dp=0.0):
    cfg = U.tiny_cfg()
    sd = O.classifier_state(cfg, C, seed=0, input_norm=inorm)
    image, audio = U.make_inputs(cfg, B)
    tw = torch.randn(B, C, generator=torch.Generator().manual_seed(5))
    gd = torch.Generator().manual_seed(9)
    drops = [(torch.rand(B, generator=gd) < 1 - dp).float() / (1 - dp) for _ in range(6 * cfg.depth)] if dp else None
    preds, stats, grads = O.classifier_loss_and_grads(sd, cfg, image, audio, tw, input_norm=inorm, training=True, freeze_encoder=freeze, drops=drops)
    model = U.build_classifier(cfg, C, freeze, inorm, device, drop_path=dp)
    assert set(model.state_dict()) == set(sd)
    model.load_state_dict(sd, strict=True)
    assert model.train() is None                   # reference quirk: AVClassifier.train() returns None (classifier.py:61-64)
    assert model.encoder.training == (not freeze)
    with U.inject_droppath(drops or []) as it:
        out = model(image.to(device), audio.to(device))
        assert next(it, None) is None                # same number of draws, same order as the reference (image, audio, fusion per layer)
    sum((p * tw.to(device)).sum() for p in out).backward()
    for p, r in zip(out, preds):
        rel = float((p.detach().cpu() - r).norm() / r.norm())
        assert rel < 2e-2, rel                      # bf16 encoder, f32 tail
    named = dict(model.named_parameters())
    assert {k for k, p in named.items() if p.requires_grad and p.grad is not None and p.grad.abs().sum() > 0} >= \
        {k for k in grads if "head" in k}
    gnorm = float(torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values())))
    for k in grads:
        if "head" in k or not freeze:
            g, r = named[k].grad.detach().cpu(), grads[k]
            # mathematically-zero gradients (e.g. the final LayerNorm bias in front of a BatchNorm) are round-off: absolute bound
            assert float((g - r).norm()) < 6e-2 * float(r.norm()) + 1e-4 * gnorm, (k, float((g - r).norm()), float(r.norm()))
    msd = model.state_dict()
    for k, v in stats.items():
        assert float((msd[k].cpu() - v).norm() / v.norm()) < 2e-2, k
    if inorm:
        assert int(msd["image_norm.num_batches_tracked"]) == 1
    # eval mode uses the running statistics (no update)
    model.eval()
    with torch.no_grad():
        e1 = model(image.to(device), audio.to(device))
    epreds, _ = O.classifier_forward({**sd, **stats}, cfg, image, audio, input_norm=inorm, training=False)
    for p, r in zip(e1, epreds):
        assert float((p.cpu() - r).norm() / r.norm()) < 3e-2
    return model


def test_gradient_accumulation_and_zero_grad(emulated):
    cfg = U.tiny_cfg()
    model = U.build_model(cfg)
    image, audio = U.make_inputs(cfg, 2)
    ni, na = U.make_noise(cfg, 2)
    def step():
        with U.inject_rand([ni, na]):
            li, la, _, _ = model(image, audio)
        (li + la).backward()
    step()
    g1 = model._davf_store.flat_g.clone()
    step()
    assert torch.allclose(model._davf_store.flat_g, 2 * g1, rtol=1e-4, atol=1e-7)      # += semantics (misc.py:144-148)
    torch.optim.SGD(model.parameters(), lr=0.1).zero_grad()                             # set_to_none=True
    assert all(p.grad is None for p in model.parameters())
    step()
    assert torch.allclose(model._davf_store.flat_g, g1, rtol=1e-4, atol=1e-7)
    assert all(p.grad is not None for p in model.parameters() if p.requires_grad)


def test_state_dict_roundtrip_and_shadow_refresh(emulated):
    cfg = U.tiny_cfg()
    model = U.build_model(cfg)
    image, audio = U.make_inputs(cfg, 1)
    ni, na = U.make_noise(cfg, 1)
    with U.inject_rand([ni, na]):
        l0 = model(image, audio)[0].item()
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    with torch.no_grad():
        for p in model.parameters():
            if p.requires_grad:
                p.mul_(1.5)
    with U.inject_rand([ni, na]):
        l1 = model(image, audio)[0].item()
    assert abs(l1 - l0) > 1e-4                       # in-place edits reach the bf16 shadows
    model.load_state_dict(sd, strict=True)
    with U.inject_rand([ni, na]):
        l2 = model(image, audio)[0].item()
    assert abs(l2 - l0) < 1e-6


def test_encoder_api_contract(emulated):
    cfg = U.tiny_cfg()
    model = U.build_model(cfg)
    enc = model.encoder
    image, audio = U.make_inputs(cfg, 2)
    xi, xa, xf, embs = enc(image, audio, return_embs=True)
    assert xi.shape == (2, 16, 128) and xa.shape == (2, 12, 128) and xf.shape == (2, 32, 128) and len(embs) == cfg.depth
    sd = {k: v for k, v in model.state_dict().items()}
    ref = O.encoder_forward(O._Prec(False), sd, cfg, image, audio)
    for a, b in zip((xi, xa, xf), ref):
        assert ((a - b).norm() / b.norm()).item() < 2e-2
    assert enc.embed_dim == 128 and enc.image.patch_embed.grid_size == (4, 4) and enc.image.patch_embed.patch_size == (16, 16)
    ids = enc.params_layer_ids()
    assert max(l for _, l in ids) == cfg.depth + 1
    # every trainable encoder parameter appears exactly once (lr_sched.py:32 builds a dict from it)
    assert {id(p) for p, _ in ids if p is not None} == {id(p) for p in enc.parameters() if p.requires_grad}


def test_unsupported_options_fail_loudly():
    from deepavfusion_b200.models import DeepAVFusion
    with pytest.raises(NotImplementedError):
        DeepAVFusion(image_pretrained="", audio_pretrained="", attn_drop=0.1)
    with pytest.raises(NotImplementedError):
        DeepAVFusion(image_pretrained="", audio_pretrained="", fusion_arch="dense_mmi")


def test_cpu_model_without_kernels_fails_loudly():
    cfg = U.tiny_cfg()
    model = U.build_model(cfg)
    image, audio = U.make_inputs(cfg, 1)
    with pytest.raises(RuntimeError, match="CUDA"):
        model(image, audio)


def test_cabi_exports_every_declared_symbol():
    """The C-ABI library loads and exports exactly what include/davf.h declares (no compute calls)."""
    from deepavfusion_b200 import _cabi
    header = open(os.path.join(ROOT, "include", "davf.h")).read()
    declared = set(re.findall(r"\b(davf_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_cabi.SIGNATURES), declared ^ set(_cabi.SIGNATURES)
    if not os.path.exists(_cabi.LIB_PATH):
        _cabi.build()
    handle = ctypes.CDLL(_cabi.LIB_PATH)
    for name in declared:
        assert hasattr(handle, name), name
    lib = _cabi.lib()
    assert lib.davf_version() == 1
    # argument validation happens before any CUDA call
    assert lib.davf_set_attn_impl(7) == -1 and b"set_attn_impl" in lib.davf_last_error()
    assert lib.davf_mask_rank(None, 1, 4, 9, None, None, None, None) == -1


def test_flop_model_matches_survey_constants():
    """bench.py's FLOPs per clip-pair (util/flops.py) reproduce the four constants SURVEY.md 8(d) derives for the BASELINE
    configurations, and the pair-factorisation discount the survey says to subtract."""
    from deepavfusion_b200.util import flops as F
    vgg, aset = F.vggsound_pretrain(), F.audioset_pretrain()
    cls = F.unmasked_classifier(310)
    assert abs(F.gflop_forward(vgg, factorised=False) - 39.01) < 0.01 and abs(F.gflop_step(vgg, factorised=False) - 116.77) < 0.01
    assert abs(F.gflop_forward(aset, factorised=False) - 43.27) < 0.01 and abs(F.gflop_step(aset, factorised=False) - 129.55) < 0.01
    assert abs(F.gflop_forward(cls, factorised=False) - 66.07) < 0.01 and abs(F.gflop_step(cls, factorised=False) - 197.94) < 0.01
    assert abs((F.gflop_step(vgg, False) - F.gflop_step(vgg, True)) - 5.94) < 0.01
    assert abs((F.gflop_step(aset, False) - F.gflop_step(aset, True)) - 9.51) < 0.01
    b = F.forward_breakdown(vgg, factorised=False)
    for k, v in dict(patch_embed=0.27, image_blocks=9.38, audio_blocks=4.17, fusion_blocks=5.88, decoder_embed=0.10,
                     image_decoder=12.33, audio_decoder=6.71, pred=0.18).items():
        assert abs(b[k] / 1e9 - v) < 0.006, (k, b[k] / 1e9, v)


def test_param_groups_lrd_mirror():
    """lr_sched.py:25-58: one (decay, no-decay) group pair per layer id with lr_scale = decay ** (top - layer)."""
    from deepavfusion_b200.util import lr_sched
    cfg = U.tiny_cfg()
    model = U.build_classifier(cfg, 10, False, False, "cpu")
    no_wd = [n for n, p in model.named_parameters() if "bias" in n or "norm" in n]
    groups = lr_sched.param_groups_lrd(model, 0.05, no_weight_decay_list=no_wd, layer_decay=0.75)
    layer_of = {id(p): l for p, l in model.params_layer_ids()}
    top = max(layer_of.values())
    seen = set()
    for g in groups:
        lids = {layer_of[id(p)] for p in g["params"]}
        assert len(lids) == 1
        lid = lids.pop()
        assert abs(g["lr_scale"] - 0.75 ** (top - lid)) < 1e-12
        for p in g["params"]:
            assert id(p) not in seen
            seen.add(id(p))
            assert (g["weight_decay"] == 0.0) == (p.ndim == 1 or any(p is q for n, q in model.named_parameters() if n in no_wd))
    assert seen == {id(p) for p in model.parameters() if p.requires_grad}
    heads = [g for g in groups if any(p is model.image_head.weight for p in g["params"])]
    assert heads and heads[0]["lr_scale"] == 1.0

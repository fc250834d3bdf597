"""-m gpu: the CUDA path (through the C ABI, on cuda:0) against the CPU oracle on identical weights,
inputs and mask noise.  Tolerances (bf16 operands, f32 accumulation / residual stream; SURVEY.md
8(c)): loss rtol 2e-3; per-tensor gradient rel-L2 <= max(3e-2, 1.5x the oracle's own bf16-autocast
error); global gradient rel-L2 <= 2e-2; mask indices bit-exact."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import avmae_oracle as O
import model_utils as U

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def run_ours(cfg, sd, image, audio, ni, na):
    model = U.build_model(cfg, "cuda")
    model.load_state_dict(sd, strict=True)
    with U.inject_rand([ni, na]):
        li, la, pi, pa = model(image.cuda(), audio.cuda())
    (li + la).backward()
    torch.cuda.synchronize()
    return model, li, la, pi, pa


@pytest.mark.parametrize("name,kw,B", [("tiny", {}, 3), ("tiny_fullwidth", dict(fusion_attn_ratio=1.0, fusion_mlp_ratio=4.0), 2),
                                       ("tiny_sparse", dict(fusion_layers="1", depth=3), 2)])
def test_tiny_fwd_bwd_vs_oracle(name, kw, B):
    cfg = U.tiny_cfg(**kw)
    sd = O.build_state(cfg, seed=0)
    image, audio = U.make_inputs(cfg, B)
    ni, na = U.make_noise(cfg, B)
    out, grads = O.loss_and_grads(sd, cfg, image, audio, ni, na)
    _, amp_grads = O.loss_and_grads(sd, cfg, image, audio, ni, na, amp=True)
    model, li, la, pi, pa = run_ours(cfg, sd, image, audio, ni, na)
    assert abs(li.item() - out["loss_image"].item()) <= 2e-3 * abs(out["loss_image"].item())
    assert abs(la.item() - out["loss_audio"].item()) <= 2e-3 * abs(out["loss_audio"].item())
    rel = lambda a, b: ((a.float().cpu() - b).norm() / b.norm()).item()
    assert rel(pi, out["pred_image"]) < 2e-2 and rel(pa, out["pred_audio"]) < 2e-2
    failures, worst, glob = U.grad_report(dict(model.named_parameters()), grads, amp_grads)
    assert not failures, f"{len(failures)} gradient tensors out of tolerance, e.g. {failures[:5]}"
    assert glob <= 2e-2, glob


@pytest.mark.parametrize("tag,freeze,inorm,dp", [("linprobe", True, True, 0.0), ("finetune", False, False, 0.0), ("finetune_bn", False, True, 0.0),
                                                 ("finetune_droppath", False, False, 0.2)])
def test_classifier_vs_oracle(tag, freeze, inorm, dp):
    """a11 / BASELINE configs 4-5 at test size: drop-in AVClassifier on cuda:0 (unmasked encoder + pool / BatchNorm /
    heads, forward and backward, then eval mode on the running statistics) against the oracle restatement."""
    from test_host_cpu import _classifier_case
    _classifier_case("cuda", freeze, inorm, dp=dp)


def test_classifier_golden_linprobe():
    """Predictions of the REAL reference classifier (tests/golden/classifier_tiny.npz) from the CUDA path."""
    z = np.load(os.path.join(GOLD, "classifier_tiny.npz"))
    cfg = U.tiny_cfg()
    image, audio = U.make_inputs(cfg, 4)
    model = U.build_classifier(cfg, 10, True, True, "cuda")
    model.load_state_dict(O.classifier_state(cfg, 10, seed=0, input_norm=True), strict=True)
    model.train()
    preds = model(image.cuda(), audio.cuda())
    for n, p in zip(("image", "audio", "fusion"), preds):
        ref = torch.from_numpy(z[f"linprobe_pred_{n}"])
        assert float((p.detach().cpu() - ref).norm() / ref.norm()) < 2e-2


def test_vitb_golden_fwd_bwd():
    """BASELINE config 1 (ViT-B, r=.25, mlp 1, B=2) against the fixture written from the REAL reference."""
    meta = json.load(open(os.path.join(GOLD, "vggsound_b2.json")))
    gold = np.load(os.path.join(GOLD, "vggsound_b2.npz"))
    cfg = O.OracleConfig(fusion_attn_ratio=0.25, fusion_mlp_ratio=1.0)
    sd = O.build_state(cfg, seed=0)
    image, audio = U.make_inputs(cfg, 2)
    ni, na = torch.from_numpy(gold["noise_image"]), torch.from_numpy(gold["noise_audio"])
    model, li, la, pi, pa = run_ours(cfg, sd, image, audio, ni, na)
    assert set(model.state_dict()) == set(meta["state_shapes"])
    assert abs(li.item() - float(gold["loss_image"])) <= 2e-3 * float(gold["loss_image"]), (li.item(), float(gold["loss_image"]))
    assert abs(la.item() - float(gold["loss_audio"])) <= 2e-3 * float(gold["loss_audio"]), (la.item(), float(gold["loss_audio"]))
    close = lambda a, b: ((a.float().cpu() - torch.from_numpy(b)).norm() / torch.from_numpy(b).norm()).item()
    assert close(pi[:, :4, :16], gold["pred_image_head"]) < 3e-2
    assert close(pa[:, :4, :16], gold["pred_audio_head"]) < 3e-2
    named = dict(model.named_parameters())
    keys = meta["grad_keys"]
    gn_ref = float(gold["grad_norm_global"])
    # (1) against the fixture of the REAL reference: per-tensor norms AND the leading elements of every gradient tensor
    # (a transposed / mis-routed gradient of the right norm fails the element check)
    norms = {k: named[k].grad.float().norm().item() for k in keys}
    gn = sum(v ** 2 for v in norms.values()) ** 0.5
    assert abs(gn - gn_ref) <= 1e-2 * gn_ref, (gn, gn_ref)
    bad = []
    for i, k in enumerate(keys):
        ref = float(gold["grad_norms"][i])
        if ref > 1e-4 * gn_ref and abs(norms[k] - ref) > 5e-2 * ref:
            bad.append((k, norms[k], ref))
        g = named[k].grad.detach().float().cpu().flatten()[:8]
        h = torch.from_numpy(gold["grad_heads"][i][:g.numel()])
        n_el = max(named[k].numel(), 1)
        tol = 0.1 * float(h.abs().max()) + 4.0 * ref / n_el ** 0.5 * 0.03 + 1e-5 * gn_ref     # bf16 error ~3 % of the tensor's rms
        if float((g - h).abs().max()) > tol:
            bad.append((k, "head", g.tolist(), h.tolist()))
    assert not bad, bad[:5]
    # (2) every gradient tensor element-wise against the oracle (itself pinned to the reference by make_golden.py)
    _, grads = O.loss_and_grads(sd, cfg, image, audio, ni, na)
    _, amp_grads = O.loss_and_grads(sd, cfg, image, audio, ni, na, amp=True)
    failures, worst, glob = U.grad_report(named, grads, amp_grads)
    assert not failures, f"{len(failures)} gradient tensors out of tolerance, e.g. {failures[:5]}"
    assert glob <= 2e-2, glob
    # encoder-only unmasked forward (AVMAE.forward_encoder; BASELINE config 4 path)
    with torch.no_grad():
        xi, xa, xf = model.forward_encoder(image.cuda(), audio.cuda())
    assert close(xf, gold["enc_x_fusion_unmasked"]) < 2e-2
    assert close(xi[:, :4, :32], gold["enc_x_image_unmasked_head"]) < 2e-2
    assert close(xa[:, :4, :32], gold["enc_x_audio_unmasked_head"]) < 2e-2


def test_mask_rng_stream_and_bit_exact_indices():
    """The noise is torch.rand on the device with the reference's draw order (image first), so for one
    seed the masks equal argsort-of-the-same-noise exactly."""
    cfg = U.tiny_cfg()
    model = U.build_model(cfg, "cuda")
    torch.manual_seed(123)
    ik, im, ir = model.random_masking(5, 16, 0.75, "cuda")
    torch.manual_seed(123)
    noise = torch.rand(5, 16, device="cuda")
    ek, em, er = O.random_masking(noise.cpu(), 0.75)
    assert torch.equal(ik.cpu(), ek) and torch.equal(im.cpu(), em) and torch.equal(ir.cpu(), er)


def test_cpu_tensor_fails_loudly():
    cfg = U.tiny_cfg()
    model = U.build_model(cfg, "cpu")
    image, audio = U.make_inputs(cfg, 1)
    with pytest.raises(RuntimeError):
        model(image, audio)


def test_multistream_backward_matches_single_stream(monkeypatch):
    """The three per-layer branches run on three streams (and the two decoders on two).  Gradients must be
    identical in value to the single-stream run, iteration after iteration -- this is the regression test for
    the uninitialised-shared-memory read (0 * NaN) that only concurrency exposed."""
    cfg = O.OracleConfig(fusion_attn_ratio=0.25, fusion_mlp_ratio=1.0)
    model = U.build_model(cfg, "cuda")
    image, audio = U.make_inputs(cfg, 8)
    image, audio = image.cuda(), audio.cuda()

    def run():
        model._davf_store.zero_grad()
        torch.manual_seed(7)
        li, la, _, _ = model(image, audio)
        (li + la).backward()
        model._davf_store.join_side_streams(torch.cuda.current_stream())
        torch.cuda.synchronize()
        return model._davf_store.flat_g.clone()

    monkeypatch.setenv("DAVF_STREAMS", "0")
    with torch.no_grad():
        model(image, audio)                       # builds the store
    ref = run()
    assert bool(torch.isfinite(ref).all())
    monkeypatch.setenv("DAVF_STREAMS", "1")
    for it in range(12):
        g = run()
        assert bool(torch.isfinite(g).all()), f"non-finite gradient in multi-stream iteration {it}"
        rel = ((g - ref).norm() / ref.norm()).item()
        assert rel < 1e-3, (it, rel)


def test_overlapped_optimizer_matches_plain_step(monkeypatch):
    """Trainer.step applies the fused AdamW bucket by bucket under backward (GradSync.fuse_optimizer); the parameters,
    optimizer state and grad norm after two steps equal those of backward-then-one-AdamW-launch (DAVF_OVERLAP_ADAMW=0)."""
    from deepavfusion_b200.util.misc import Trainer
    cfg = U.tiny_cfg()
    sd = O.build_state(cfg, seed=0)
    image, audio = U.make_inputs(cfg, 3)
    ni, na = U.make_noise(cfg, 3)
    res = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("DAVF_OVERLAP_ADAMW", mode)
        model = U.build_model(cfg, "cuda")
        model.load_state_dict(sd, strict=True)
        tr = Trainer(model, optimizer=torch.optim.AdamW(model.parameters(), lr=1e-3, betas=(0.9, 0.95), weight_decay=0.05), bucket_mb=0.25)
        assert (tr.sync is not None and tr.sync.optimizer is not None and len(tr.sync.buckets) > 3) == (mode == "1")
        for _ in range(2):
            with U.inject_rand([ni, na]):
                li, la, _, _ = model(image.cuda(), audio.cuda())
            norm, _ = tr.step(li + la)
        torch.cuda.synchronize()
        assert float(tr.store.flat_g.abs().max()) == 0.0          # zero_grad is part of the fused step
        res[mode] = (tr.store.flat_p.clone(), tr.optimizer.flat_m.clone(), tr.optimizer.flat_v.clone(), float(norm), tr.optimizer.n_steps)
    p1, m1, v1, n1, s1 = res["1"]
    p0, m0, v0, n0, s0 = res["0"]
    assert s1 == s0 == 2 and abs(n1 - n0) <= 1e-3 * n0
    p_init = torch.cat([sd[k].flatten() for k in sd]).norm()
    assert float((p1 - p0).norm()) <= 1e-5 * float(p0.norm())
    assert float((m1 - m0).norm()) <= 1e-3 * float(m0.norm()) and float((v1 - v0).norm()) <= 1e-3 * float(v0.norm())


def test_graph_step_async_host_inputs_match_device_inputs():
    """GraphedTrainStep.step_async (pinned host inputs over the copy stream, losses read back through the pinned ring
    one step late) produces the same per-step losses / grad norms as the plain call with device-resident inputs."""
    from deepavfusion_b200.util.misc import Trainer
    from deepavfusion_b200.util.graphed import GraphedTrainStep
    cfg = U.tiny_cfg()
    sd = O.build_state(cfg, seed=0)
    image, audio = U.make_inputs(cfg, 3)
    h_img, h_aud = image.pin_memory(), audio.pin_memory()
    seqs = {}
    for mode in ("device", "async"):
        model = U.build_model(cfg, "cuda")
        model.load_state_dict(sd, strict=True)
        tr = Trainer(model, optimizer=torch.optim.AdamW(model.parameters(), lr=1e-3, betas=(0.9, 0.95)))
        torch.manual_seed(11)
        g = GraphedTrainStep(tr, image.cuda(), audio.cuda(), warmup=1)
        model.load_state_dict(sd, strict=True)                    # warm-up steps moved the weights: start both modes equal
        tr.optimizer.flat_m.zero_(); tr.optimizer.flat_v.zero_()
        torch.manual_seed(12)
        out = []
        for _ in range(3):
            if mode == "device":
                li, la, n = g(image.cuda(), audio.cuda())
                out.append((float(li), float(la), float(n)))
            else:
                g.step_async(h_img, h_aud)
                if g.pending() > 1:
                    out.append(g.pop_metrics())
        while mode == "async" and g.pending():
            out.append(g.pop_metrics())
        seqs[mode] = out
        del g
    assert len(seqs["async"]) == 3
    for a, b in zip(seqs["device"], seqs["async"]):
        for x, y in zip(a, b):
            assert abs(x - y) <= 2e-3 * abs(x) + 1e-6, (seqs["device"], seqs["async"])


def test_graph_capture_is_side_effect_free_and_accumulates_like_eager():
    """GraphedTrainStep (a) leaves parameters, Adam moments, beta^t and the step counters exactly as it found them (its
    warm-up steps and capture pass must not train) and (b) with accum_iter = 2 -- two graphs: accumulate-only micro-step
    and final micro-step with the fused AdamW, misc.py:144-148 -- walks the same trajectory as the eager Trainer."""
    from deepavfusion_b200.util.misc import Trainer
    from deepavfusion_b200.util.graphed import GraphedTrainStep
    cfg = U.tiny_cfg()
    sd = O.build_state(cfg, seed=0)
    B, accum, steps = 3, 2, 2
    micro = [tuple(t.cuda() for t in U.make_inputs(cfg, B, seed=20 + i)) for i in range(accum)]
    ni, na = (t.cuda() for t in U.make_noise(cfg, B))
    static = {tuple(ni.shape): ni, tuple(na.shape): na}
    orig_rand = torch.rand

    def fake_rand(*size, **kw):                      # the same mask noise in every step, eager and captured (device-side clone)
        return static[tuple(size)].clone()
    torch.rand = fake_rand
    try:
        runs = {}
        for mode in ("eager", "graph"):
            model = U.build_model(cfg, "cuda")
            model.load_state_dict(sd, strict=True)
            tr = Trainer(model, optimizer=torch.optim.AdamW(model.parameters(), lr=1e-3, betas=(0.9, 0.95), weight_decay=0.05), accum_iter=accum)
            norms = []
            if mode == "graph":
                opt = tr.optimizer
                before = [t.clone() for t in (tr.store.flat_p, tr.store.flat_lp, tr.store.flat_g, opt.flat_m, opt.flat_v, opt.scal)]
                g = GraphedTrainStep(tr, *micro[0], warmup=2)
                after = (tr.store.flat_p, tr.store.flat_lp, tr.store.flat_g, opt.flat_m, opt.flat_v, opt.scal)
                for b, a in zip(before, after):
                    assert torch.equal(b, a), "graph construction changed training state"
                assert opt.n_steps == 0 and int(tr.n_steps) == 0 and tr.accums == 0
                assert g.graph_micro is not None and g.launches_micro > 0 and g.launches_final > g.launches_micro
            for _ in range(steps):
                for m in range(accum):
                    if mode == "graph":
                        out = g(*micro[m])
                        norm = out[-1]
                        assert (norm is None) == (m < accum - 1)
                    else:
                        li, la, _, _ = model(*micro[m])
                        norm, _ = tr.step(li + la)
                    if norm is not None:
                        norms.append(float(norm))
            torch.cuda.synchronize()
            assert tr.optimizer.n_steps == steps and int(tr.n_steps) == steps
            runs[mode] = (tr.store.flat_p.clone(), tr.optimizer.flat_m.clone(), tr.optimizer.flat_v.clone(), norms)
            if mode == "graph":
                del g
    finally:
        torch.rand = orig_rand
    (pe, me, ve, ne), (pg, mg, vg, ng) = runs["eager"], runs["graph"]
    assert len(ne) == len(ng) == steps
    for a, b in zip(ne, ng):
        assert abs(a - b) <= 2e-3 * abs(a), (ne, ng)
    assert float((pe - pg).norm()) <= 1e-5 * float(pe.norm())
    assert float((me - mg).norm()) <= 2e-3 * float(me.norm()) and float((ve - vg).norm()) <= 2e-3 * float(ve.norm())


def test_checkpoint_manager_resume_gpu(tmp_path):
    """f1 on the CUDA path: CheckpointManager file layout, torch.optim.AdamW-shaped optimizer entry, resume == uninterrupted."""
    from test_trainer_cpu import _checkpoint_resume_case
    _checkpoint_resume_case("cuda", tmp_path)


def test_return_embs_matches_oracle_layers():
    """f3 (deepavfusion.py:108-109, knn_probe.py:99): ``return_embs=True`` hands back every layer's (x_image, x_audio,
    x_fusion); the last entry, final-normed, is the encoder output, and every layer matches the oracle's activations."""
    cfg = U.tiny_cfg()
    sd = O.build_state(cfg, seed=0)
    image, audio = U.make_inputs(cfg, 2)
    model = U.build_model(cfg, "cuda")
    model.load_state_dict(sd, strict=True)
    with torch.no_grad():
        xi, xa, xf, embs = model.encoder(image.cuda(), audio.cuda(), return_embs=True)
        yi, ya, yf = model.encoder(image.cuda(), audio.cuda())
    assert len(embs) == cfg.depth and all(len(e) == 3 for e in embs)
    assert torch.equal(xi, yi) and torch.equal(xa, ya) and torch.equal(xf, yf)
    P = O._Prec(False)
    ri, ra, rf = O.encoder_forward(P, sd, cfg, image, audio)
    rel = lambda a, b: float((a.float().cpu() - b).norm() / b.norm())
    assert rel(xi, ri) < 2e-2 and rel(xa, ra) < 2e-2 and rel(xf, rf) < 2e-2
    for e in embs:
        assert e[0].shape == xi.shape and e[1].shape == xa.shape and e[2].shape == xf.shape and bool(torch.isfinite(e[2]).all())

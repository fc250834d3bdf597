"""CPU: the oracle against the fixtures written from the REAL reference (oracle/make_golden.py)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import avmae_oracle as O
import model_utils as U

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _digest(t):
    import hashlib
    return hashlib.sha256(t.detach().contiguous().numpy().tobytes()).hexdigest()[:16]


@pytest.mark.parametrize("name,cfg,B,nparams", [
    ("vggsound_b2", O.OracleConfig(fusion_attn_ratio=0.25, fusion_mlp_ratio=1.0), 2, 320_563_712),
    ("audioset_b1", O.OracleConfig(fusion_attn_ratio=1.0, fusion_mlp_ratio=4.0), 1, 378_997_760),
])
def test_oracle_matches_reference_fixture(name, cfg, B, nparams):
    meta = json.load(open(os.path.join(GOLD, f"{name}.json")))
    gold = np.load(os.path.join(GOLD, f"{name}.npz"))
    shapes = O.state_shapes(cfg)
    assert {k: list(v) for k, v in shapes.items()} == meta["state_shapes"]          # 893-key state_dict contract
    assert sum(int(np.prod(s)) for s in shapes.values()) == nparams == meta["n_params"]   # checkpoint-size KAT
    sd = O.build_state(cfg, seed=0)
    image, audio = U.make_inputs(cfg, B)
    # the recipe must regenerate the exact tensors the fixture was made from
    assert _digest(image) == meta["digests"]["image"] and _digest(audio) == meta["digests"]["audio"]
    assert _digest(torch.cat([sd[k].flatten()[:64] for k in sorted(sd)])) == meta["digests"]["state"]
    ni, na = torch.from_numpy(gold["noise_image"]), torch.from_numpy(gold["noise_audio"])
    out, grads = O.loss_and_grads(sd, cfg, image, audio, ni, na)
    assert torch.equal(out["image_ids_keep"], torch.from_numpy(gold["image_ids_keep"]))
    assert torch.equal(out["image_ids_restore"], torch.from_numpy(gold["image_ids_restore"]))
    assert torch.equal(out["image_mask"], torch.from_numpy(gold["image_mask"]))
    assert torch.equal(out["audio_ids_keep"], torch.from_numpy(gold["audio_ids_keep"]))
    assert abs(out["loss_image"].item() - float(gold["loss_image"])) < 1e-5 * float(gold["loss_image"])
    assert abs(out["loss_audio"].item() - float(gold["loss_audio"])) < 1e-5 * float(gold["loss_audio"])
    np.testing.assert_allclose(out["pred_image"][:, :4, :16].numpy(), gold["pred_image_head"], rtol=1e-3, atol=1e-4)
    keys = meta["grad_keys"]
    assert set(keys) == set(grads)
    norms = np.array([grads[k].double().norm().item() for k in keys])
    big = gold["grad_norms"] > 1e-6
    np.testing.assert_allclose(norms[big], gold["grad_norms"][big], rtol=2e-3)
    heads = np.stack([np.pad(grads[k].flatten()[:8].numpy(), (0, max(0, 8 - grads[k].numel()))) for k in keys])
    np.testing.assert_allclose(heads, gold["grad_heads"], rtol=5e-2, atol=1e-5 * float(gold["grad_norm_global"]))
    # self-consistency identities
    assert float(out["image_mask"].sum()) == B * (196 - 49) and float(out["audio_mask"].sum()) == B * (96 - 19)
    for k in keys:                                     # zero K-bias gradients (softmax shift invariance)
        if k.endswith("attn.k.bias"):
            assert grads[k].norm().item() < 1e-4 * float(gold["grad_norm_global"])


def test_mask_ties_fixture():
    d = np.load(os.path.join(GOLD, "mask_ties.npz"))
    k, m, r = O.random_masking(torch.from_numpy(d["noise"]), 0.75)
    assert (k.numpy() == d["ids_keep"]).all() and (r.numpy() == d["ids_restore"]).all() and (m.numpy() == d["mask"]).all()
    # inverse-permutation identity
    assert torch.equal(torch.gather(r, 1, k), torch.arange(k.shape[1]).expand(k.shape[0], -1))


def test_pair_factorisation_identity():
    """softmax over the 8x8 (v,a) pairs == outer product of two 8-way softmaxes."""
    torch.manual_seed(0)
    D, qk, nv, na = 64, 16, 8, 8
    v, a, q = torch.randn(nv, D).double(), torch.randn(na, D).double(), torch.randn(5, qk).double()
    Wk, bk, Wv, bv = torch.randn(qk, 2 * D).double(), torch.randn(qk).double(), torch.randn(D, 2 * D).double(), torch.randn(D).double()
    xva = torch.cat([v[:, None].expand(-1, na, -1), a[None].expand(nv, -1, -1)], -1).flatten(0, 1)
    ref = torch.softmax(q @ (xva @ Wk.t() + bk).t() * 0.125, -1) @ (xva @ Wv.t() + bv)
    pv = torch.softmax(q @ (v @ Wk[:, :D].t() + bk).t() * 0.125, -1)
    pa = torch.softmax(q @ (a @ Wk[:, D:].t()).t() * 0.125, -1)
    fac = pv @ (v @ Wv[:, :D].t() + bv) + pa @ (a @ Wv[:, D:].t())
    assert (ref - fac).abs().max().item() < 1e-12


def test_adamw_restatement_matches_torch():
    torch.manual_seed(0)
    p = torch.randn(1000)
    g = torch.randn(1000) * 0.1
    ref = torch.nn.Parameter(p.clone())
    opt = torch.optim.AdamW([ref], lr=1e-3, betas=(0.9, 0.95), weight_decay=0.05)
    m, v, mine = torch.zeros(1000), torch.zeros(1000), p.clone()
    for step in (1, 2, 3):
        ref.grad = g.clone()
        opt.step()
        O.adamw_step(mine, g, m, v, step, 1e-3, 0.9, 0.95, 1e-8, 0.05)
    assert (mine - ref.detach()).abs().max().item() < 1e-6


def test_classifier_oracle_matches_reference_fixture():
    """a11: oracle AVClassifier restatement vs outputs of the real reference classifier (tests/golden/classifier_tiny.npz,
    written by oracle/make_golden.py): predictions, gradient norms and BatchNorm running statistics."""
    import model_utils as U
    z = np.load(os.path.join(GOLD, "classifier_tiny.npz"))
    cfg = U.tiny_cfg()
    C, B = 10, 4
    image, audio = U.make_inputs(cfg, B)
    tw = torch.from_numpy(z["target_w"])
    drops = [torch.from_numpy(r) for r in z["droppath_scales"]]          # the masks the reference's DropPath modules were fed
    for tag, freeze, inorm in (("linprobe", True, True), ("finetune", False, False), ("finetune_droppath", False, False)):
        sd = O.classifier_state(cfg, C, seed=0, input_norm=inorm)
        preds, stats, grads = O.classifier_loss_and_grads(sd, cfg, image, audio, tw, input_norm=inorm, training=True, freeze_encoder=freeze,
                                                          drops=drops if tag.endswith("droppath") else None)
        for n, p in zip(("image", "audio", "fusion"), preds):
            ref = torch.from_numpy(z[f"{tag}_pred_{n}"])
            assert float((p - ref).norm() / ref.norm()) < 1e-5
        keys = [str(k) for k in z[f"{tag}_grad_keys"]]
        assert sorted(grads) == keys
        mine = np.array([float(grads[k].double().norm()) for k in keys])
        np.testing.assert_allclose(mine, z[f"{tag}_grad_norms"], rtol=1e-4, atol=1e-7)
        for k, v in stats.items():
            np.testing.assert_allclose(v.numpy(), z[f"{tag}_{k}"], rtol=1e-5, atol=1e-7)


def test_logmel_oracle_matches_torchaudio_fixture():
    """The front-end oracle (oracle/logmel_oracle.py) against the fixture written from the reference's own torchaudio
    pipeline (train.py:50-54, datasets.py:242) by oracle/make_golden_logmel.py."""
    from oracle import logmel_oracle as L
    z = np.load(os.path.join(GOLD, "logmel.npz"))
    wave = L.make_wave()
    assert np.array_equal(wave[:, :64].numpy(), z["wave_head"])
    out = L.log_mel(wave, torch.from_numpy(z["gain_db"]))
    ref = torch.from_numpy(z["logmel"])
    assert out.shape == ref.shape == (3, 1, 128, 192)
    assert float((out - ref).abs().max()) < 2e-3          # torchaudio evaluates in f32: its quiet bands carry ~1e-3 of rounding
    assert float((out - ref).abs()[ref > ref.amax(dim=2, keepdim=True) - 4.0].max()) < 3e-4      # within 40 dB of the frame's loudest band

"""CPU: Trainer / FusedAdamW / GradSync host logic with emulated kernels; the N>1 path runs as two
gloo processes (world_size 2) and must reproduce the single-process full-batch update."""
import os
import socket
import sys
from types import SimpleNamespace

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import cpu_kernels
import model_utils as U
from oracle import avmae_oracle as O


class _Opt(dict):
    __getattr__ = dict.__getitem__


def _groups(model):
    from deepavfusion_b200.util import lr_sched
    no_wd = [n for n, p in model.named_parameters() if "bias" in n or "norm" in n]        # train.py:88
    return lr_sched.param_groups_pretrained(model, 0.05, no_weight_decay_list=no_wd, image_pt="x", audio_pt=None)


def _train_steps(model, image, audio, noises, steps, accum=1, distributed=False):
    from deepavfusion_b200.util.misc import Trainer
    opt = torch.optim.AdamW(_groups(model), lr=1e-3, betas=(0.9, 0.95))
    for gi, g in enumerate(opt.param_groups):
        g["lr"] = 1e-3 * (1 + gi)                      # distinct per-group LRs must be honoured
    trainer = Trainer(model, optimizer=opt, accum_iter=accum, distributed=distributed)
    norms = []
    for s in range(steps * accum):
        with U.inject_rand(list(noises)):
            li, la, _, _ = trainer.model(image, audio)
        norm, _ = trainer.step(li + la)
        if norm is not None:
            norms.append(float(norm))
    return trainer, norms


def test_fused_adamw_matches_torch_adamw(monkeypatch):
    cpu_kernels.install(monkeypatch)
    cfg = U.tiny_cfg()
    image, audio = U.make_inputs(cfg, 2)
    noises = U.make_noise(cfg, 2)
    sd = O.build_state(cfg, seed=0)
    ours = U.build_model(cfg); ours.load_state_dict(sd)
    trainer, norms = _train_steps(ours, image, audio, noises, steps=3)
    # reference loop: same model code, stock torch AdamW on the same groups
    ref = U.build_model(cfg); ref.load_state_dict(sd)
    opt = torch.optim.AdamW(_groups(ref), lr=1e-3, betas=(0.9, 0.95))
    for gi, g in enumerate(opt.param_groups):
        g["lr"] = 1e-3 * (1 + gi)
    ref_norms = []
    for s in range(3):
        with U.inject_rand(list(noises)):
            li, la, _, _ = ref(image, audio)
        opt.zero_grad()
        (li + la).backward()
        ref_norms.append(float(torch.sqrt(sum((p.grad.double() ** 2).sum() for p in ref.parameters() if p.grad is not None))))
        opt.step()
    # Adam is chaotic in the elements whose gradient is ~eps (the update is lr * sign-like), and a 1-ulp
    # f32 difference after step 1 can flip a bf16 weight rounding: compare step 1 tightly (separate test
    # below) and the 3-step trajectory in norm.
    num = sum(((a - b) ** 2).sum() for (_, a), (_, b) in zip(ours.named_parameters(), ref.named_parameters())).sqrt().item()
    den = sum(((b - sd[n]) ** 2).sum() for n, b in ref.named_parameters()).sqrt().item()
    assert num <= 0.1 * den, (num, den)
    assert all(abs(a - b) <= 1e-2 * b for a, b in zip(norms, ref_norms)), (norms, ref_norms)
    assert float(trainer.store.flat_g.abs().max()) == 0            # zero_grad fused into the step
    # frozen pos-embeds untouched, bf16 shadows refreshed by the step
    assert torch.equal(ours.encoder.image.pos_embed, ref.encoder.image.pos_embed)
    assert torch.equal(trainer.store.flat_lp, trainer.store.flat_p.to(torch.bfloat16))
    # optimizer checkpoint entry is torch.optim.AdamW-shaped and round-trips
    sd_opt = trainer.optimizer.state_dict()
    ref_sd = opt.state_dict()
    assert set(sd_opt["state"]) == set(ref_sd["state"]) and sd_opt["state"][0].keys() == ref_sd["state"][0].keys()
    for k in ref_sd["state"]:
        assert sd_opt["state"][k]["exp_avg"].shape == ref_sd["state"][k]["exp_avg"].shape
        assert float(sd_opt["state"][k]["step"]) == float(ref_sd["state"][k]["step"]) == 3
    trainer.optimizer.load_state_dict(ref_sd)
    assert trainer.optimizer.n_steps == 3


def test_fused_adamw_single_step_tight(monkeypatch):
    cpu_kernels.install(monkeypatch)
    cfg = U.tiny_cfg()
    image, audio = U.make_inputs(cfg, 2)
    noises = U.make_noise(cfg, 2)
    sd = O.build_state(cfg, seed=0)
    ours = U.build_model(cfg); ours.load_state_dict(sd)
    _train_steps(ours, image, audio, noises, steps=1)
    ref = U.build_model(cfg); ref.load_state_dict(sd)
    opt = torch.optim.AdamW(_groups(ref), lr=1e-3, betas=(0.9, 0.95))
    for gi, g in enumerate(opt.param_groups):
        g["lr"] = 1e-3 * (1 + gi)
    with U.inject_rand(list(noises)):
        li, la, _, _ = ref(image, audio)
    opt.zero_grad()
    (li + la).backward()
    opt.step()
    for (n, a), (_, b) in zip(ours.named_parameters(), ref.named_parameters()):
        assert torch.allclose(a, b, rtol=1e-5, atol=2e-6), n


def test_gradient_accumulation_equals_big_batch(monkeypatch):
    cpu_kernels.install(monkeypatch)
    cfg = U.tiny_cfg()
    image, audio = U.make_inputs(cfg, 2)
    noises = U.make_noise(cfg, 2)
    sd = O.build_state(cfg, seed=0)
    a = U.build_model(cfg); a.load_state_dict(sd)
    b = U.build_model(cfg); b.load_state_dict(sd)
    _train_steps(a, image, audio, noises, steps=1, accum=1)
    _train_steps(b, image, audio, noises, steps=1, accum=2)        # same micro-batch twice, /2 folded into AdamW
    for (n, x), (_, y) in zip(a.named_parameters(), b.named_parameters()):
        assert torch.allclose(x, y, rtol=1e-4, atol=1e-6), n


def test_lr_schedule_mirror():
    from deepavfusion_b200.util import lr_sched
    args = SimpleNamespace(opt=_Opt(lr=1e-3, epochs=300, warmup_epochs=50, pt_warmup_epochs="300/2", pt_lr_mult_start=0, pt_lr_mult_end=1))
    opt = SimpleNamespace(param_groups=[{"lr": 0}, {"lr": 0, "pretrained": True}, {"lr": 0, "lr_scale": 0.5}])
    assert abs(lr_sched.adjust_learning_rate(opt, 25, args) - 5e-4) < 1e-12
    assert opt.param_groups[1]["lr"] < opt.param_groups[0]["lr"] and abs(opt.param_groups[2]["lr"] - 2.5e-4) < 1e-12
    lr = lr_sched.adjust_learning_rate(opt, 175, args)
    assert abs(lr - 1e-3 * 0.5) < 1e-9 and opt.param_groups[1]["lr"] == lr


def test_bucketed_optimizer_step_equals_plain_step(monkeypatch):
    """Trainer.step with the optimizer applied bucket by bucket from inside backward (GradSync.fuse_optimizer; forced on for
    CPU tensors here) leaves exactly the parameters / moments / grad norm of backward followed by one whole-buffer step."""
    cpu_kernels.install(monkeypatch)
    from deepavfusion_b200.util.misc import Trainer
    cfg = U.tiny_cfg()
    image, audio = U.make_inputs(cfg, 2)
    ni, na = U.make_noise(cfg, 2)
    res = {}
    for mode in ("force", "0"):
        monkeypatch.setenv("DAVF_OVERLAP_ADAMW", mode)
        model = U.build_model(cfg); model.load_state_dict(O.build_state(cfg, seed=0))
        tr = Trainer(model, optimizer=torch.optim.AdamW(_groups(model), lr=1e-3, betas=(0.9, 0.95)), accum_iter=2, bucket_mb=0.25)
        assert (tr.sync is not None and tr.sync.optimizer is not None) == (mode == "force")
        for it in range(4):                              # two optimizer steps of two micro-steps each
            with U.inject_rand([ni, na]):
                li, la, _, _ = model(image, audio)
            norm, _ = tr.step(li + la)
            assert (norm is None) == (it % 2 == 0)
            if it % 2 == 0:
                assert float(tr.store.flat_g.abs().max()) > 0        # accumulated, not yet consumed
        assert float(tr.store.flat_g.abs().max()) == 0.0 and tr.optimizer.n_steps == 2
        res[mode] = (tr.store.flat_p.clone(), tr.optimizer.flat_m.clone(), tr.optimizer.flat_v.clone(), float(norm))
    for a, b in zip(res["force"][:3], res["0"][:3]):
        assert torch.equal(a, b)
    assert abs(res["force"][3] - res["0"][3]) <= 1e-6 * res["0"][3]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _dp_worker(rank, world, port, out_dir, fused=False):
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                      DAVF_OVERLAP_ADAMW="force" if fused else "0")
    torch.set_num_threads(2)
    cpu_kernels.install()
    from deepavfusion_b200.util.misc import Trainer
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg = U.tiny_cfg()
    image, audio = U.make_inputs(cfg, 2 * world)
    ni, na = U.make_noise(cfg, 2 * world)
    sl = slice(2 * rank, 2 * rank + 2)
    model = U.build_model(cfg); model.load_state_dict(O.build_state(cfg, seed=rank))   # rank-dependent init: broadcast must fix it
    opt = torch.optim.AdamW(_groups(model), lr=1e-3, betas=(0.9, 0.95))
    trainer = Trainer(model, optimizer=opt, accum_iter=2, distributed=True, bucket_mb=0.25)
    assert trainer.sync is not None and len(trainer.sync.buckets) > 3
    for micro in range(2):
        with U.inject_rand([ni[sl], na[sl]]):
            li, la, _, _ = trainer.model(image[sl], audio[sl])
        if fused:                                        # all-reduce + bucketed AdamW from inside backward (Trainer.step)
            trainer.step(li + la)
        else:
            trainer.backward(li + la)
        if micro == 0:                                   # no_sync semantics: nothing reduced yet
            assert not any(trainer.sync.launched)
    grads = trainer.store.flat_g.clone() * float(trainer.optimizer.scal[2])
    if fused:
        assert float(grads.abs().max()) == 0.0 and trainer.optimizer.n_steps == 1
    else:
        trainer.optimizer.step()
    torch.save({"grads": grads, "names": trainer.store.names, "offsets": trainer.store.offsets,
                "state": {k: v.clone() for k, v in model.state_dict().items()}}, os.path.join(out_dir, f"rank{rank}{'f' if fused else ''}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_data_parallel_gloo_world2(tmp_path, monkeypatch):
    world = 2
    mp.spawn(_dp_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r0, r1 = torch.load(tmp_path / "rank0.pt"), torch.load(tmp_path / "rank1.pt")
    assert torch.equal(r0["grads"], r1["grads"])
    # the same step with the all-reduce followed bucket by bucket by the fused AdamW (Trainer.step): identical parameters
    mp.spawn(_dp_worker, args=(world, _free_port(), str(tmp_path), True), nprocs=world, join=True)
    f0, f1 = torch.load(tmp_path / "rank0f.pt"), torch.load(tmp_path / "rank1f.pt")
    for k in r0["state"]:
        assert torch.equal(f0["state"][k], f1["state"][k]) and torch.equal(f0["state"][k], r0["state"][k]), f"fused step differs on {k}"
    for k in r0["state"]:
        assert torch.equal(r0["state"][k], r1["state"][k]), f"ranks diverged on {k}"
    # single process, full batch: the averaged 2-rank gradient equals the full-batch gradient
    cpu_kernels.install(monkeypatch)
    from deepavfusion_b200.util.misc import Trainer
    cfg = U.tiny_cfg()
    image, audio = U.make_inputs(cfg, 4)
    noises = U.make_noise(cfg, 4)
    model = U.build_model(cfg); model.load_state_dict(O.build_state(cfg, seed=0))
    trainer = Trainer(model, optimizer=torch.optim.AdamW(_groups(model), lr=1e-3, betas=(0.9, 0.95)), accum_iter=2)
    for micro in range(2):
        with U.inject_rand(list(noises)):
            li, la, _, _ = model(image, audio)
        trainer.backward(li + la)
    full = trainer.store.flat_g * float(trainer.optimizer.scal[2])
    assert r0["names"] == trainer.store.names
    rel = ((full - r0["grads"]).norm() / full.norm()).item()
    # not 1e-6: the emulated GEMMs (MKL f32) are not bit-identical per row for different batch sizes, and a
    # 1-ulp f32 difference occasionally flips a bf16 rounding; the same split in ONE process gives 1.4e-3.
    assert rel < 5e-3, rel


def _checkpoint_resume_case(device, tmp_path):
    """a run checkpointed through CheckpointManager (reference layout, misc.py:222-309) and resumed into
    fresh objects continues exactly like the uninterrupted run; the 'optimizer' entry loads into a stock
    ``torch.optim.AdamW`` (reference train.py:93) and a state written by the stock optimizer loads back."""
    from deepavfusion_b200.util.misc import CheckpointManager, Trainer
    cfg = U.tiny_cfg()
    image, audio = (t.to(device) for t in U.make_inputs(cfg, 2))
    noises = [n.to(device) for n in U.make_noise(cfg, 2)]
    sd = O.build_state(cfg, seed=0)

    def fresh():
        model = U.build_model(cfg, device); model.load_state_dict(sd)
        opt = torch.optim.AdamW(_groups(model), lr=1e-3, betas=(0.9, 0.95))
        return Trainer(model, optimizer=opt)

    def run(trainer, steps):
        for _ in range(steps):
            with U.inject_rand([n.clone() for n in noises]):
                li, la, _, _ = trainer.model(image, audio)
            trainer.step(li + la)

    full = fresh(); run(full, 4)                                   # uninterrupted
    first = fresh(); run(first, 2)
    mgr = CheckpointManager(first.module_dict(), str(tmp_path), epochs=10, save_freq=1)
    mgr.checkpoint(1, {"epoch": 1, "best_loss": 0.5}, is_best=True)
    for name in ("checkpoint_latest.pth", "checkpoint_best.pth", "checkpoint_0001.pth"):
        assert (tmp_path / name).is_file()
    ckpt = torch.load(tmp_path / "checkpoint_latest.pth", map_location="cpu", weights_only=False)
    assert set(ckpt) == {"state_dict", "optimizer", "n_steps", "epoch", "best_loss"} and int(ckpt["n_steps"]) == 2
    assert set(ckpt["state_dict"]) == set(sd) and all(not v.is_cuda for v in ckpt["state_dict"].values())

    # the optimizer entry is torch.optim.AdamW's: a stock optimizer over same-shaped parameters loads it ...
    clones = [[torch.nn.Parameter(p.detach().cpu().clone()) for p in g["params"]] for g in first.optimizer.param_groups]
    stock = torch.optim.AdamW([{"params": ps} for ps in clones], lr=1e-3, betas=(0.9, 0.95))
    stock.load_state_dict(ckpt["optimizer"])
    st0 = stock.state[clones[1][0]]
    ours0 = first.optimizer.state[first.optimizer.param_groups[1]["params"][0]]
    assert float(st0["step"]) == 2 and torch.equal(st0["exp_avg"], ours0["exp_avg"].cpu()) and torch.equal(st0["exp_avg_sq"], ours0["exp_avg_sq"].cpu())
    # ... and what the stock optimizer writes loads back
    first.optimizer.load_state_dict(stock.state_dict())
    assert first.optimizer.n_steps == 2

    resumed = fresh()
    mgr2 = CheckpointManager(resumed.module_dict(), str(tmp_path), epochs=10)
    epoch, metrics = mgr2.resume()
    assert epoch == 1 and metrics == {"best_loss": 0.5} and int(resumed.n_steps) == 2 and resumed.optimizer.n_steps == 2
    # state right after the resume is bit-identical to the run that wrote the file
    assert torch.equal(resumed.store.flat_p, first.store.flat_p)
    assert torch.equal(resumed.optimizer.flat_m, first.optimizer.flat_m) and torch.equal(resumed.optimizer.flat_v, first.optimizer.flat_v)
    assert torch.equal(resumed.optimizer.scal[:2].cpu(), first.optimizer.scal[:2].cpu())
    run(resumed, 2)
    # and it continues like the uninterrupted run: the only differences are beta^t rebuilt as b ** t instead of t f32
    # multiplications (1 ulp) and, on the GPU, the order of the f32 atomics in the weight gradients
    p0 = fresh().store.flat_p
    num, den = float((resumed.store.flat_p - full.store.flat_p).norm()), float((full.store.flat_p - p0).norm())
    assert num <= 2e-3 * den, (num, den)
    assert int(resumed.n_steps) == int(full.n_steps) == 4 and resumed.optimizer.n_steps == 4


def test_checkpoint_manager_resume(monkeypatch, tmp_path):
    cpu_kernels.install(monkeypatch)
    _checkpoint_resume_case("cpu", tmp_path)


def test_rehome_grads_keeps_training_identical(monkeypatch):
    """ParamStore.rehome_grads (used to move flat_g into NCCL-registered memory) re-points every gradient view: accumulated
    gradients survive the move and the following steps equal those of a store that never moved."""
    cpu_kernels.install(monkeypatch)
    from deepavfusion_b200.util.misc import Trainer
    cfg = U.tiny_cfg()
    image, audio = U.make_inputs(cfg, 2)
    ni, na = U.make_noise(cfg, 2)
    res = {}
    for move in (False, True):
        model = U.build_model(cfg); model.load_state_dict(O.build_state(cfg, seed=0))
        tr = Trainer(model, optimizer=torch.optim.AdamW(_groups(model), lr=1e-3, betas=(0.9, 0.95)), accum_iter=2)
        for it in range(4):
            with U.inject_rand([ni, na]):
                li, la, _, _ = model(image, audio)
            norm, _ = tr.step(li + la)
            if move and it == 0:                          # mid-accumulation: the half-summed gradients must come along
                old = tr.store.flat_g
                tr.store.rehome_grads(torch.empty_like(old))
                assert tr.store.flat_g.data_ptr() != old.data_ptr() and torch.equal(tr.store.flat_g, old)
                for k, p in enumerate(tr.store.params):
                    if p.requires_grad:
                        assert p.grad.data_ptr() == tr.store.flat_g[tr.store.offsets[k]:].data_ptr()
                old.fill_(float("nan"))                   # nothing may read the old buffer any more
        res[move] = (tr.store.flat_p.clone(), float(norm))
    assert torch.equal(res[True][0], res[False][0]) and res[True][1] == res[False][1]


def test_bucket_layout_and_registration_fallback(monkeypatch):
    """GradSync buckets tile [0, numel) from the end on ALIGN boundaries for any bucket / tail size, and the NCCL-registered
    gradient buffer quietly reports why it is unavailable (off by default; no NCCL process group on CPU)."""
    cpu_kernels.install(monkeypatch)
    from deepavfusion_b200.models.layers import ensure_store
    from deepavfusion_b200.params import ALIGN
    from deepavfusion_b200.util import distributed as D
    model = U.build_model(U.tiny_cfg())
    store = ensure_store(model)
    for bucket_mb, tail_mb in ((0.25, None), (0.25, 0.03), (64.0, 16.0), (0.05, 0.05)):
        sync = D.GradSync(store, bucket_mb=bucket_mb, tail_mb=tail_mb)
        ranges = sorted(sync.ranges)
        assert ranges[0][0] == 0 and max(hi for _, hi in ranges) == store.numel
        for (lo, hi), (lo2, _) in zip(ranges, ranges[1:]):
            assert hi <= lo2 and lo % ALIGN == 0 and lo2 % ALIGN == 0       # disjoint, on the grid (gaps are alignment padding)
        assert sorted(k for ks in sync.buckets for k in ks) == list(range(len(store.params)))
        assert sync.ranges[0][1] == store.numel                              # bucket 0 is the END of the buffer (first out of backward)
    fine = D.GradSync(store, bucket_mb=0.25, tail_mb=0.03)
    coarse = D.GradSync(store, bucket_mb=0.25)
    assert len(fine.buckets) > len(coarse.buckets)
    store.sync = None
    buf, why = D.nccl_registered_zeros(1024, "cpu")
    assert buf is None and "DAVF_NCCL_REGISTER" in why
    monkeypatch.setenv("DAVF_NCCL_REGISTER", "1")
    buf, why = D.nccl_registered_zeros(1024, "cpu")
    assert buf is None and "NCCL" in why

#!/usr/bin/env python
"""Benchmark of the DeepAVFusion hot path on B200 (contract: see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config vggsound|audioset|featex|finetune]
                    [--impl ours|reference] [--no-graph]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Default workload (the one the driver runs) = BASELINE.json configs[1], "vggsound": one optimizer step of the VGGSound
pre-training recipe (ViT-B, fusion attn_ratio 0.25 / mlp_ratio 1.0, bf16, batch 64 per GPU) on synthetic inputs: mask
draw, forward, backward, gradient all-reduce (N > 1) and the fused AdamW update.  The other BASELINE configurations:
  audioset   configs[2]: attn_ratio 1.0 / mlp_ratio 4.0, batch 64 per GPU, accum_iter 4 (README.md:51-55); a step =
             4 micro-steps (two CUDA graphs: accumulate-only, and final with all-reduce + AdamW; misc.py:144-148)
  featex     configs[3]: frozen-encoder feature extraction, AVClassifier.eval() forward, no masking, batch 256
             (eval_linprobe.py:90-102)
  finetune   configs[4]: AVClassifier fwd + bwd, unmasked, batch 32, accum_iter 4, drop_path 0.2, layer-wise lr decay
             (configs/finetune.yaml:36-50, eval_finetune.py:161-214)
Rank 0 prints ONE JSON line.
  value     AV clip-pairs/s, whole job, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e       same metric through the public API with pinned HOST inputs: H2D copies and a D2H read of the step's
            result inside the timed region, every step
  roofline  the tcgen05 GEMM kernel (dominant kernel): FLOPs of the step's GEMM launches / their measured duration,
            against the measured cuBLAS bf16 peaks in MEASURED_PEAKS.json (burst for the isolated replay)
  cpu_baseline      the CPU oracle (oracle/avmae_oracle.py, a restatement of the reference) timed on the host
            cores on a bounded sample (BASELINE.json configs[0]: batch 2, fp32)
  torch_eager_bf16  the same oracle run on the GPU through eager PyTorch ops under bf16 autocast semantics at the
            workload's batch size -- what the reference does on this box, the practical bar
  dp        (N > 1) rank-equality of the parameters after the timed region, DP-vs-full-batch gradient error,
            and what NCCL used for the exchange step
--impl reference times the CPU arm alone, on all host threads (rank 0 only under torchrun).
"""
from __future__ import annotations

import argparse
import contextlib
import glob
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "AV clip-pairs/sec (fwd+bwd, ViT-B pretrain)"
# name -> workload description.  FLOPs per clip-pair come from deepavfusion_b200/util/flops.py (SURVEY.md 8(d)
# convention, pair factorisation subtracted as the survey requires; pinned to the survey's constants by a CPU test).
CONFIGS = {
    "vggsound": dict(workload="vggsound_pretrain_vitb_fusion-all_r0.25_mlp1.0_bf16_b64-per-gpu", kind="pretrain", r=0.25, mlp=1.0,
                     batch=64, accum=1, step="mask + fwd + bwd + grad all-reduce + fused AdamW"),
    "audioset": dict(workload="audioset_pretrain_vitb_fusion-all_r1.0_mlp4.0_bf16_b64-per-gpu_accum4", kind="pretrain", r=1.0, mlp=4.0,
                     batch=64, accum=4, step="4 x (mask + fwd + bwd) + grad all-reduce + fused AdamW"),
    "featex": dict(workload="feature_extraction_vitb_unmasked_fwd_bf16_b256", kind="featex", r=0.25, mlp=1.0,
                   batch=256, accum=1, step="AVClassifier.eval() forward, no masking, all tokens"),
    "finetune": dict(workload="finetune_vitb_unmasked_fwd+bwd_bf16_b32_accum4_droppath0.2", kind="finetune", r=0.25, mlp=1.0,
                     batch=32, accum=4, step="4 x (unmasked fwd + bwd, drop_path 0.2) + grad all-reduce + fused AdamW (layer decay 0.75)"),
}
NUM_CLASSES = 310


def gflop_per_pair(cfg, attention=True):
    from deepavfusion_b200.util import flops as F
    if cfg["kind"] == "pretrain":
        return F.gflop_step(F.PathShape(fusion_attn_ratio=cfg["r"], fusion_mlp_ratio=cfg["mlp"]), attention=attention)
    s = F.unmasked_classifier(NUM_CLASSES, cfg["r"], cfg["mlp"])
    return F.gflop_forward(s, attention=attention) if cfg["kind"] == "featex" else F.gflop_step(s, attention=attention)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(tflops=float(d["bf16_tflops_sustained"]), tflops_burst=float(d["bf16_tflops"]), hbm=float(d["hbm_gbs"]),
                    src="measured (MEASURED_PEAKS.json, cuBLAS bf16)")
    return dict(tflops=1590.0, tflops_burst=1590.0, hbm=6650.0, src="fallback (B200_PROFILING.md)")


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_arm(steps: int, warmup: int, batch: int = 2, r: float = 0.25, mlp: float = 1.0):
    import torch
    from oracle import avmae_oracle as O
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    cfg = O.OracleConfig(fusion_attn_ratio=r, fusion_mlp_ratio=mlp)
    sd = O.build_state(cfg, seed=0)
    g = torch.Generator().manual_seed(1)
    image = torch.randn(batch, 3, 224, 224, generator=g)
    audio = torch.randn(batch, 1, 128, 192, generator=g)
    ni, na = torch.rand(batch, 196, generator=g), torch.rand(batch, 96, generator=g)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        O.loss_and_grads(sd, cfg, image, audio, ni, na)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    med = statistics.median(times)
    return dict(value=batch / med, unit="clip-pairs/s", cores=cores, kind="port",
                sample=f"oracle fwd+bwd, ViT-B r{r}/mlp{mlp}, batch {batch}, fp32, {steps} timed iterations (median {med:.3f} s)"), med


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = CONFIGS[args.config]
    steps, warmup = max(1, min(args.steps, 20)), max(1, min(args.warmup, 3))
    base, med = cpu_arm(steps, warmup, r=cfg["r"], mlp=cfg["mlp"])
    line = dict(impl="reference", metric=METRIC, value=base["value"], unit="clip-pairs/s",
                n_gpus=args.gpus, steps=steps, warmup=warmup, ms_per_step=med * 1e3, higher_is_better=True,
                scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
                config=dict(workload=cfg["workload"], note="CPU arm: bounded sample of the same model at batch 2 per step (BASELINE configs[0]), "
                            "pre-training fwd+bwd of the oracle port on all host threads"),
                cpu_baseline=base,
                e2e=dict(value=base["value"], unit="clip-pairs/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# clocks sampling during the timed region
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock / throttle reasons of one GPU sampled every 50 ms during the timed region: NVML in a thread
    (nvidia-ml-py), falling back to an `nvidia-smi -lms 100` subprocess."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.idx, self.stop = [], None, gpu_index, False

    def _nvml_loop(self):
        import pynvml as N
        h = N.nvmlDeviceGetHandleByIndex(self.idx)
        mx = N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM)
        bits = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))
        while not self.stop:
            try:
                r = N.nvmlDeviceGetCurrentClocksEventReasons(h) if hasattr(N, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else N.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.rows.append([str(N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM)), str(mx), "0"] +
                                 ["Active" if (r & b) else "Not Active" for _, b in bits])
            except Exception:
                pass
            time.sleep(0.05)

    def __enter__(self):
        try:
            import pynvml as N
            N.nvmlInit()
            # NVML indexes physical devices; honour CUDA_VISIBLE_DEVICES when it is a plain index list
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            if vis and all(x.strip().isdigit() for x in vis.split(",")):
                self.idx = int(vis.split(",")[self.idx])
            self.t = threading.Thread(target=self._nvml_loop, daemon=True)
            self.t.start()
            return self
        except Exception:
            pass
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.idx)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *a):
        self.stop = True
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def mark(self):
        """Rows sampled before this call (sampler start-up, warm-up steps) are not part of the timed region."""
        self.start = len(self.rows)

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        for r in self.rows[getattr(self, "start", 0):]:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        load = [x for x in sm if x > 0.5 * mx] or sm
        return dict(sm_mhz=statistics.median(load) if load else None, sm_max_mhz=mx or None, reasons=sorted(reasons), samples=len(sm))


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def build_encoder(cfg, drop_path=0.0):
    from deepavfusion_b200.models import DeepAVFusion
    return DeepAVFusion(image_arch="vit_base", image_pretrained="", image_size=(224, 224),
                        audio_arch="vit_base", audio_pretrained="", audio_size=(128, 192),
                        fusion_arch="factorized_mmi", fusion_layers="all", num_fusion_tkns=(16, 8, 8),
                        fusion_mlp_ratio=cfg["mlp"], fusion_attn_ratio=cfg["r"], fusion_num_heads=12, drop_path=drop_path)


def build_trainer(cfg, device, distributed):
    """The model / optimizer / Trainer exactly as train.py:66-103 (pre-training) or eval_finetune.py:161-214 builds them."""
    import torch
    from deepavfusion_b200.models import AVMAE, AVClassifier
    from deepavfusion_b200.util import lr_sched
    from deepavfusion_b200.util.misc import Trainer
    torch.manual_seed(0)
    world = max(1, int(os.environ.get("WORLD_SIZE", "1")))
    if cfg["kind"] == "pretrain":
        enc = build_encoder(cfg)
        model = AVMAE(enc, enc.embed_dim, image_decoder_depth=8, image_mask_ratio=0.75, image_norm_loss=True,
                      audio_decoder_depth=8, audio_mask_ratio=0.8, audio_norm_loss=True).to(device)
        no_wd = [n for n, p in model.named_parameters() if "bias" in n or "norm" in n]              # train.py:88
        groups = lr_sched.param_groups_pretrained(model, 0.05, no_weight_decay_list=no_wd, image_pt="vit_base_mae_in1k", audio_pt="vit_base_audiomae_as2m")
        lr = 1.5e-4 * cfg["batch"] * cfg["accum"] * world / 256                                      # train.py:32-35
        opt = torch.optim.AdamW(groups, lr=lr, betas=(0.9, 0.95))                                    # train.py:93
    else:
        model = AVClassifier(build_encoder(cfg, drop_path=0.2), NUM_CLASSES, freeze_encoder=False, input_norm=False).to(device)
        no_wd = [n for n, p in model.named_parameters() if "bias" in n or "norm" in n]
        groups = lr_sched.param_groups_lrd(model, 0.05, no_weight_decay_list=no_wd, layer_decay=0.75)   # eval_finetune.py:199-203
        lr = 3e-4 * cfg["batch"] * cfg["accum"] * world / 256
        for g in groups:
            g["lr"] = lr * g["lr_scale"]                                                           # lr_sched.py:21-23
        opt = torch.optim.AdamW(groups, lr=lr)
        model.train()
    return Trainer(model, optimizer=opt, accum_iter=cfg["accum"], use_amp=True, distributed=distributed)


def finetune_loss(model, image, audio, target):
    """eval_finetune.py:285-293 with joint_loss (finetune.yaml:33) and soft targets (mixup / label smoothing)."""
    import torch
    pi, pa, pf = model(image, audio)
    preds = (pi + pa + pf) / 3.0
    loss = torch.sum(-target * torch.log_softmax(preds.float(), dim=-1), dim=-1).mean()
    return loss, (loss.detach(),)


def synth_inputs(cfg, batch, seed, pinned):
    import torch
    g = torch.Generator().manual_seed(seed)
    image = torch.randn(batch, 3, 224, 224, generator=g)
    audio = (1.5 * torch.randn(batch, 1, 128, 192, generator=g) - 3).clamp_(-7, 3)               # log-mel-like
    out = [image, audio]
    if cfg["kind"] == "finetune":
        out.append(torch.softmax(4.0 * torch.randn(batch, NUM_CLASSES, generator=g), dim=-1))       # soft targets
    if pinned:
        out = [t.pin_memory() for t in out]
    return out


@contextlib.contextmanager
def inject_rand(noises):
    """Make the next torch.rand(N, L, device=...) draws (AVMAE.random_masking, avmae.py:127) return the given noise."""
    import torch
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


def gemm_roofline(trace, peaks, reps=5):
    """Replay the step's GEMM launch list (same shapes / operand layouts / epilogues) back to back and time
    it with CUDA events on the launching stream."""
    import torch
    import deepavfusion_b200.kernels as K
    bf16 = torch.bfloat16
    cache = {}

    def buf(key, shape, dtype):
        k = (key, tuple(shape), dtype)
        if k not in cache:
            cache[k] = (torch.randn(*shape, device="cuda") * 0.05).to(dtype) if dtype != torch.int64 else torch.zeros(*shape, dtype=dtype, device="cuda")
        return cache[k]
    calls = []
    flops = 0.0

    def problem(t):
        nonlocal flops
        M, N, Kd = t["M"], t["N"], t["K"]
        a = buf("a", (M if t["a_kmajor"] else Kd, t["lda"]), bf16)[:, :(Kd if t["a_kmajor"] else M)]
        b = buf("b", (N if t["b_kmajor"] else Kd, t["ldb"]), bf16)[:, :(Kd if t["b_kmajor"] else N)]
        out = buf("o", (M, t["ldo"]), bf16 if t["out_bf16"] else torch.float32)[:, :N]
        kw = dict(bias=buf("bias", (N,), torch.float32) if t["bias"] else None, act=t["act"], want_aux=t["aux"],
                  res=buf("r", (M, N), torch.float32) if t["res"] else None,
                  res_idx=buf("ri", (M,), torch.int64) if t["res_idx"] else None,
                  out=out, accumulate=t["accumulate"],
                  rowsum_out=buf("rs", (M,), torch.float32) if t.get("rowsum") else None)
        if t["act"] == K.ACT_DGELU:
            kw["aux_in"] = buf("h", (M, N), bf16)
        flops += 2.0 * M * N * Kd
        return (a, b, t["a_kmajor"], t["b_kmajor"]), kw
    for t in trace:              # one entry per LAUNCH: a single problem or a grouped launch of several
        calls.append([problem(u) for u in t["group"]] if "group" in t else problem(t))

    def replay():
        for c in calls:
            if isinstance(c, list):
                K.gemm_grouped(c)
            else:
                K.gemm(*c[0], **c[1])
    replay()                     # warm-up (instantiates kernels, TMA descriptors)
    torch.cuda.synchronize()
    # The launch list is captured into a CUDA graph, exactly as the training step runs it: issued from Python the
    # ~700 launches cost 10-30 us of host time each, which would time the interpreter instead of the kernels.
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, capture_error_mode="thread_local"):     # (an NCCL watchdog thread may be alive in this process)
        replay()
    graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        graph.replay()
    e1.record()
    torch.cuda.synchronize()
    sec = e0.elapsed_time(e1) / 1e3 / reps
    achieved = flops / sec / 1e12
    # The replay is a short isolated burst (tens of ms at the boost clock): its denominator is the BURST cuBLAS figure.
    # ``frac_sustained`` (same numerator over the sustained figure) is kept for comparison with round 1.
    return dict(bound="tensor", achieved=achieved, peak=peaks["tflops_burst"], unit="TFLOP/s", frac=achieved / peaks["tflops_burst"],
                frac_sustained=achieved / peaks["tflops"], traffic=None,
                kernel="davf::gemm_tc_kernel (tcgen05/TMA)", launches_per_step=len(calls), gemm_ms_per_step=sec * 1e3,
                flops_per_launch=flops / len(calls), avg_launch_us=sec * 1e6 / len(calls),
                peak_source=peaks["src"] + "; burst for the isolated replay, sustained for the in-step figure"), sec



def eager_bf16_leg(cfg, batch, steps=5, warmup=2):
    """The practical bar on the same box: the reference's graph executed by eager
    PyTorch on this GPU under bf16-autocast semantics -- the oracle restatement with amp=True on cuda tensors (cuBLAS
    GEMMs, SDPA attention in the ViT blocks as timm's fused_attn does, ATen elementwise), fwd + bwd + torch.optim.AdamW."""
    import torch
    from oracle import avmae_oracle as O
    dev = torch.device("cuda")
    ocfg = O.OracleConfig(fusion_attn_ratio=cfg["r"], fusion_mlp_ratio=cfg["mlp"])
    sd = {k: v.to(dev) for k, v in O.build_state(ocfg, seed=0).items()}
    leaves = {k: v.requires_grad_(k not in O.FROZEN_KEYS) for k, v in sd.items()}
    opt = torch.optim.AdamW([v for v in leaves.values() if v.requires_grad], lr=1e-4, betas=(0.9, 0.95), weight_decay=0.05)
    g = torch.Generator(device=dev).manual_seed(1)
    image = torch.randn(batch, 3, 224, 224, generator=g, device=dev)
    audio = torch.randn(batch, 1, 128, 192, generator=g, device=dev)
    times = []
    for i in range(warmup + steps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ni, na = torch.rand(batch, 196, device=dev), torch.rand(batch, 96, device=dev)
        out = O.avmae_forward(leaves, ocfg, image, audio, ni, na, amp=True, sdpa=True)
        (out["loss_image"] + out["loss_audio"]).backward()
        opt.step()
        opt.zero_grad(set_to_none=True)
        e1.record()
        torch.cuda.synchronize()
        if i >= warmup:
            times.append(e0.elapsed_time(e1))
    med = statistics.median(times)
    del opt, leaves, sd
    torch.cuda.empty_cache()
    return dict(value=batch / (med / 1e3), unit="clip-pairs/s", ms_per_step=med, batch=batch,
                what="oracle restatement of the reference on cuda:0, eager PyTorch ops, bf16 autocast semantics (cuBLAS + SDPA), "
                     "fwd + bwd + torch.optim.AdamW, median of %d steps" % steps)


def nccl_summary():
    """What NCCL used for the exchange step, from its own INFO log (NCCL_DEBUG_FILE set in main())."""
    pat = os.environ.get("DAVF_NCCL_LOG_GLOB")
    if not pat:
        return None
    lines = []
    for fn in sorted(glob.glob(pat)):
        try:
            lines += open(fn, errors="replace").read().splitlines()
        except OSError:
            pass
    keep = [l.split("NCCL INFO", 1)[-1].strip() for l in lines if "NCCL INFO" in l]
    pick = lambda *keys: [l for l in keep if any(k in l for k in keys)]
    ver = pick("NCCL version")
    chans = pick("coll channels", "collnet channels", "nvls channels", " channels, ")
    nvls = pick("NVLS")
    algos = [l for l in pick("Algo", "algo") if "->" in l]          # "AllReduce: N Bytes -> Algo RING proto SIMPLE channel{Lo..Hi}={0..31}"
    return dict(version=ver[:1], channels=chans[:2], nvls=bool(nvls), nvls_lines=nvls[:2], tuning=sorted(set(algos))[:6], log_lines=len(keep))


def dp_checks(trainer, cfg, dev, rank, world):
    """Driver-visible proof of the DDP contract (misc.py:32-34,144-148): (1) every rank holds bit-identical parameters
    after the timed steps; (2) the bucketed, all-reduced gradient of a sharded batch equals the gradient of the same
    global batch computed on one rank."""
    import torch
    import torch.distributed as dist
    st = trainer.store
    # (1) parameter checksums: f64 sum, f64 sum of squares and a bit-pattern xor-fold, all-gathered
    p = st.flat_p
    bits = p.view(torch.int32)
    chk = torch.stack([p.double().sum(), (p.double() ** 2).sum(), bits.sum(dtype=torch.int64).double()])
    allc = [torch.empty_like(chk) for _ in range(world)]
    dist.all_gather(allc, chk)
    identical = all(bool(torch.equal(c, allc[0])) for c in allc)
    out = dict(params_identical_across_ranks=identical, param_checksum=[float(x) for x in allc[0]])
    if cfg["kind"] != "pretrain":
        return out
    # (2) DP gradient vs full-batch gradient (8 pairs per rank, the same noise on both sides)
    b = 8
    g = torch.Generator().manual_seed(4242)
    G = b * world
    image = torch.randn(G, 3, 224, 224, generator=g)
    audio = torch.randn(G, 1, 128, 192, generator=g)
    ni, na = torch.rand(G, 196, generator=g), torch.rand(G, 96, generator=g)
    sl = slice(rank * b, (rank + 1) * b)
    sync = trainer.sync
    trainer.zero_grad()
    sync.reset()
    sync.enabled, sync.fuse_optimizer = True, False
    with inject_rand([ni[sl], na[sl]]):
        li, la, _, _ = trainer.model(image[sl].to(dev), audio[sl].to(dev))
    (li + la).backward()
    st.join_side_streams(torch.cuda.current_stream())
    sync.finish()                                   # bucketed NCCL all-reduce (sum)
    torch.cuda.synchronize()
    g_dp = (st.flat_g / world).clone()
    trainer.zero_grad()
    if rank == 0:
        sync.enabled = False
        with inject_rand([ni, na]):
            li, la, _, _ = trainer.model(image.to(dev), audio.to(dev))
        (li + la).backward()
        st.join_side_streams(torch.cuda.current_stream())
        torch.cuda.synchronize()
        g_full = st.flat_g
        out["dp_grad_rel_err"] = float((g_dp - g_full).norm() / g_full.norm())
        out["dp_grad_check"] = f"{b} pairs per rank x {world} ranks through GradSync buckets vs the same {G} pairs on rank 0"
        trainer.zero_grad()
    sync.reset()
    dist.barrier()
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    import deepavfusion_b200.kernels as K
    from deepavfusion_b200.util import distributed as dist_utils
    from deepavfusion_b200.util.graphed import GraphedTrainStep

    cfg = CONFIGS[args.config]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world > 1:
        # a multi-rank run that stops making progress (a collective some rank never joins) must end on its own with the
        # Python stacks on stderr instead of sitting there until the caller's limit: a normal run takes under a minute
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ.get("DAVF_BENCH_WATCHDOG_S", "600")), exit=True)
    local = dist_utils.init_from_env("nccl") if world > 1 else 0
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    assert args.gpus == world, f"--gpus {args.gpus} but WORLD_SIZE={world} (launch with torchrun for N > 1)"
    peaks = measured_peaks()
    B, accum, kind = cfg["batch"], cfg["accum"], cfg["kind"]
    pairs_per_step = B * accum * world
    use_graph = not args.no_graph

    h_inputs = synth_inputs(cfg, B, 1000 + rank, pinned=True)
    d_inputs = [t.to(dev) for t in h_inputs]
    torch.manual_seed(2000 + rank)

    if kind == "featex":
        from deepavfusion_b200.models import AVClassifier
        torch.manual_seed(0)
        model = AVClassifier(build_encoder(cfg), NUM_CLASSES, freeze_encoder=True, input_norm=True).to(dev)
        model.eval()
        trainer = None
    else:
        trainer = build_trainer(cfg, dev, distributed=world > 1)
        model = trainer.model
    loss_fn = finetune_loss if kind == "finetune" else None

    def eager_micro(inputs, final):
        if kind == "finetune":
            loss, metrics = finetune_loss(model, *inputs)
        else:
            li, la, _, _ = model(*inputs)
            loss, metrics = li + la, (li.detach(), la.detach())
        norm, _ = trainer.step(loss)
        return (*metrics, norm)

    def eager_step(inputs):
        if kind == "featex":
            with torch.no_grad():
                return model(*inputs)
        out = None
        for m in range(accum):
            out = eager_micro(inputs, m == accum - 1)
        return out

    # one traced eager step: GEMM launch list + launch count per step
    K.GEMM_TRACE = []
    n0 = K.launch_count()
    eager_step(d_inputs)
    torch.cuda.synchronize()
    launches_per_step = K.launch_count() - n0
    trace, K.GEMM_TRACE = K.GEMM_TRACE, None

    if args.profile_step:                     # for `ncu --profile-from-start off`: exactly one eager step is profiled
        eager_step(d_inputs)
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
        eager_step(d_inputs)
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
        return

    gstep = None
    if use_graph and kind == "featex":
        # static-input forward graph (no optimizer state): inputs copied in, predictions read out
        static_in = [torch.empty_like(t) for t in d_inputs]
        for a, b_ in zip(static_in, d_inputs):
            a.copy_(b_)
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s), torch.no_grad():
            for _ in range(2):
                model(*static_in)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        fgraph = torch.cuda.CUDAGraph()
        n0 = K.launch_count()
        with torch.cuda.graph(fgraph, capture_error_mode="thread_local"), torch.no_grad():
            static_out = model(*static_in)
        launches_per_step = K.launch_count() - n0
        cstream = torch.cuda.Stream()

        def step_fn(inputs):
            if inputs[0].device.type == "cpu":
                with torch.cuda.stream(cstream):
                    cstream.wait_stream(torch.cuda.current_stream())
                    for a, b_ in zip(static_in, inputs):
                        a.copy_(b_, non_blocking=True)
                torch.cuda.current_stream().wait_stream(cstream)
            else:
                for a, b_ in zip(static_in, inputs):
                    a.copy_(b_, non_blocking=True)
            fgraph.replay()
            return static_out
    elif use_graph:
        gstep = GraphedTrainStep(trainer, *d_inputs, warmup=2, loss_fn=loss_fn)
        launches_per_step = gstep.launches_per_step

        def step_fn(inputs):
            out = None
            for _ in range(accum):
                out = gstep(*inputs)
            return out
    else:
        step_fn = eager_step

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, per_step=False):
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)] if per_step else None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for i in range(steps):
            if per_step:
                evs[i].record()
            fn()
        if per_step:
            evs[steps].record()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        each = [evs[i].elapsed_time(evs[i + 1]) for i in range(steps)] if per_step else None
        return ms, each

    # ---- device-resident throughput --------------------------------------------------------------
    last = {}

    def dev_step():
        last["out"] = step_fn(d_inputs)
    W = max(3, args.warmup)
    with ClockSampler(local) as cs:                 # nvidia-smi needs ~0.3 s to start: launched before the warm-up steps,
        for _ in range(W):                          # only rows sampled during the timed region are summarised
            dev_step()
        torch.cuda.synchronize()
        cs.mark()
        ms, each = timed(dev_step, args.steps, per_step=True)
    clocks = cs.summary()
    out = last["out"]
    result_val = float(sum(float(t.float().sum()) if t.numel() > 1 else float(t) for t in out if t is not None)) if kind == "featex" \
        else float(sum(float(t) for t in out[:-1]))
    assert result_val == result_val and abs(result_val) < 1e9, f"non-finite result {result_val}"
    ms_per_step = ms / args.steps
    value = pairs_per_step / (ms_per_step / 1e3)

    # ---- end to end: pinned host inputs -> H2D -> step -> D2H result, every step ---------------------
    sink = []
    if kind == "featex":
        host_out = [torch.empty(t.shape, dtype=t.dtype).pin_memory() for t in out]

        def e2e_step():
            res = step_fn(h_inputs)
            for h, t in zip(host_out, res):
                h.copy_(t, non_blocking=True)          # the features / predictions ARE the result of this workload
            torch.cuda.current_stream().synchronize()
            sink.append(float(host_out[0][0, 0]))

        def e2e_drain():
            pass
        d2h = sum(t.numel() * t.element_size() for t in out)
        readback = "pinned D2H of the three prediction tensors every step"
    elif use_graph:
        def e2e_step():
            # pinned host inputs -> H2D -> graph replays -> async D2H of the metrics; the metrics of the PREVIOUS micro-steps
            # are read on the host while this one runs (every step's loss is read, one step late: train.py:166 logs it)
            for _ in range(accum):
                gstep.step_async(*h_inputs)
                if gstep.pending() > 1:
                    sink.append(gstep.pop_metrics()[0])

        def e2e_drain():
            while gstep.pending() > 0:
                sink.append(gstep.pop_metrics()[0])
        d2h = 4 * len(out) * accum
        readback = "pinned async D2H of (losses..., grad_norm) every micro-step, read on the host one step later"
    else:
        def e2e_step():
            res = step_fn([t.to(dev, non_blocking=True) for t in h_inputs])
            sink.append(float(res[0]))                  # D2H read of the step's loss (train.py:166)

        def e2e_drain():
            pass
        d2h = 4
        readback = "loss.item() every step"
    for _ in range(3):
        e2e_step()
    e2e_drain()
    n_sink = len(sink)

    def e2e_run(n):                                     # n steps launched AND their results read inside the timed region
        for _ in range(n):
            e2e_step()
        e2e_drain()
    e2e_ms = timed(lambda: e2e_run(args.steps), 1)[0] / args.steps
    assert len(sink) - n_sink >= args.steps and all(v == v for v in sink[n_sink:]), "e2e results missing or NaN"
    e2e_value = pairs_per_step / (e2e_ms / 1e3)
    h2d = sum(t.numel() * t.element_size() for t in h_inputs) * accum

    # ---- data-parallel correctness, visible to the driver ---------------------------------------------
    dp = None
    if world > 1 and trainer is not None:
        dp = dp_checks(trainer, cfg, dev, rank, world)
        if rank == 0:
            dp["nccl"] = nccl_summary()
            dp["gemm_sms_backward"] = 148 - int(os.environ.get("DAVF_COMM_SMS", "32"))      # forward launches use all 148
            dp["grad_buffer_nccl_registered"] = trainer.grad_buffer_registered               # True, or why not
            dp["nccl_high_priority_stream"] = os.environ.get("DAVF_NCCL_HIGH_PRIORITY", "0") == "1"

    # ---- roofline of the dominant kernel + baselines (rank 0) -----------------------------------------
    roof, base, eager = None, None, None
    gf = gflop_per_pair(cfg)
    if rank == 0:
        roof, gemm_sec = gemm_roofline(trace, peaks)
        roof["share_of_step"] = roof["gemm_ms_per_step"] / ms_per_step
        # the launches also compute q for the fusion-token rows the reference discards (deepavfusion.py:104-105): those FLOPs
        # are executed but not algorithmic, so the roofline numerator is the survey's GEMM-only count, not sum 2MNK
        roof["achieved_executed"] = roof["achieved"]
        roof["achieved"] = gflop_per_pair(cfg, attention=False) * B * accum / 1e3 / gemm_sec
        roof["frac"] = roof["achieved"] / roof["peak"]
        roof["frac_sustained"] = roof["achieved"] / peaks["tflops"]
        roof["flops_per_launch"] = gflop_per_pair(cfg, attention=False) * B * accum * 1e9 / roof["launches_per_step"]
        step_tflops = value / world * gf / 1e3
        roof["in_step"] = dict(achieved=step_tflops, peak=peaks["tflops"], unit="TFLOP/s", frac=step_tflops / peaks["tflops"],
                               what="ALL algorithmic FLOPs of the step / step time, per GPU, against the sustained cuBLAS figure")
        prof = os.path.join(ROOT, "profiles", "gemm_traffic.json")
        if os.path.exists(prof):
            try:
                roof["traffic"] = json.load(open(prof)).get("dram_bytes_per_launch")
            except Exception:
                pass
    if rank == 0 and world == 1 and not args.skip_cpu:
        base, _ = cpu_arm(steps=5, warmup=1, r=cfg["r"], mlp=cfg["mlp"])
        if kind == "pretrain" and not args.skip_eager:
            try:
                eager = eager_bf16_leg(cfg, B)
            except Exception as e:                      # reported, never fatal for the product measurement
                eager = dict(error=f"{type(e).__name__}: {e}"[:300])

    if rank == 0:
        line = dict(metric=METRIC, value=value, unit="clip-pairs/s", n_gpus=world,
                    steps=args.steps, warmup=W, ms_per_step=ms_per_step, ms_per_step_median=statistics.median(each),
                    higher_is_better=True, scaling="weak", vs_baseline=None, dtype="bf16", data="synthetic",
                    config=dict(workload=cfg["workload"], global_batch=pairs_per_step, batch_per_gpu=B, accum_iter=accum, parallelism=f"dp{world}",
                                step=cfg["step"], cuda_graph=bool(use_graph),
                                streams=os.environ.get("DAVF_STREAMS", "1") != "0",
                                allreduce=("none" if world == 1 or trainer is None else (("bucketed NCCL all-reduce captured inside the step graph, overlapped with backward" if getattr(gstep, "overlap_comm", False)
                                                                                          else "one NCCL call after the captured fwd+bwd graph") if use_graph else "bucketed, overlapped with backward")),
                                l2="per-step working set (640 MB bf16 weights + >4 GB activations) exceeds the 126 MB L2; no flush needed",
                                gflop_per_pair=gf, step_tflops=value * gf / 1e3, step_frac_of_peak=value * gf / 1e3 / (peaks["tflops"] * world)),
                    e2e=dict(value=e2e_value, unit="clip-pairs/s", h2d_bytes_per_step=h2d * world, d2h_bytes_per_step=d2h * world,
                             readback=readback, ms_per_step=e2e_ms),
                    gpu_launches=int(launches_per_step * args.steps), launches_per_step=int(launches_per_step),
                    clocks=clocks, roofline=roof, loss=result_val)
        if base is not None:
            line["cpu_baseline"] = base
        if eager is not None:
            line["torch_eager_bf16"] = eager
        if dp is not None:
            line["dp"] = dp
        print(json.dumps(line), flush=True)
    if world > 1:
        # No NCCL teardown: destroying the process group while a captured graph still holds NCCL kernel nodes hung at
        # exit on this stack.  Every rank has synchronised its device; leave without running destructors.
        torch.cuda.synchronize()
        sys.stdout.flush(); sys.stderr.flush()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--config", default="vggsound", choices=sorted(CONFIGS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of the whole-step CUDA graph")
    ap.add_argument("--skip-cpu", action="store_true", help="skip the cpu_baseline and torch_eager_bf16 legs")
    ap.add_argument("--skip-eager", action="store_true", help="skip the torch_eager_bf16 leg")
    ap.add_argument("--profile-step", action="store_true", help="run one eager step between cudaProfilerStart/Stop and exit (for ncu)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    if int(os.environ.get("WORLD_SIZE", "1")) > 1 and "NCCL_DEBUG_FILE" not in os.environ:
        # record what NCCL picked for the exchange step (algorithm / protocol / channels / NVLS) in its own log files
        d = os.path.join("/tmp", f"davf_nccl_{os.environ.get('MASTER_PORT', '0')}")
        os.makedirs(d, exist_ok=True)
        os.environ["NCCL_DEBUG"] = "INFO"
        os.environ["NCCL_DEBUG_SUBSYS"] = "INIT,TUNING"
        os.environ["NCCL_DEBUG_FILE"] = os.path.join(d, "rank%h_%p.log")
        os.environ["DAVF_NCCL_LOG_GLOB"] = os.path.join(d, "rank*.log")
    run_ours(args)


if __name__ == "__main__":
    main()

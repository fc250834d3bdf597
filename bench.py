#!/usr/bin/env python
"""Benchmark of the DeepAVFusion pre-training hot path on B200 (contract: see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--no-graph]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one optimizer step of BASELINE.json configs[1] (VGGSound pre-train: ViT-B, fusion
attn_ratio 0.25 / mlp_ratio 1.0, bf16, batch 64 per GPU) on synthetic inputs: mask draw, forward,
backward, gradient all-reduce (N > 1) and the fused AdamW update.  Rank 0 prints ONE JSON line.
  value     AV clip-pairs/s, whole job, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e       same metric through the public Trainer API with pinned HOST inputs: H2D copies and a D2H read
            of the loss inside the timed region, every step
  roofline  the tcgen05 GEMM kernel (dominant kernel): algorithmic FLOPs of the step's GEMM launches /
            their measured duration, against the measured cuBLAS bf16 peak in MEASURED_PEAKS.json
  cpu_baseline  the CPU oracle (oracle/avmae_oracle.py, a restatement of the reference) timed on the host
            cores on a bounded sample (BASELINE.json configs[0]: batch 2, fp32)
--impl reference times that CPU arm alone, on all host threads (rank 0 only under torchrun).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = "vggsound_pretrain_vitb_fusion-all_r0.25_mlp1.0_bf16_b64-per-gpu"
BATCH_PER_GPU = 64
# Algorithmic FLOPs per clip-pair for this config: F_step = 3 F_fwd - F_patch_embed
# = 116.77 GF with dead fusion rows excluded; the build also applies the pair factorisation, so
# 3 * (2.265 - 0.283) = 5.94 GF are subtracted as the survey requires.
GFLOP_PER_PAIR = 116.77 - 5.94


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(tflops=float(d["bf16_tflops_sustained"]), hbm=float(d["hbm_gbs"]), src="measured (MEASURED_PEAKS.json, sustained cuBLAS bf16)")
    return dict(tflops=1590.0, hbm=6650.0, src="fallback (B200_PROFILING.md)")


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_arm(steps: int, warmup: int, batch: int = 2):
    import torch
    from oracle import avmae_oracle as O
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    cfg = O.OracleConfig(fusion_attn_ratio=0.25, fusion_mlp_ratio=1.0)
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
                sample=f"oracle fwd+bwd, ViT-B r0.25/mlp1, batch {batch}, fp32, {steps} timed iterations (median {med:.3f} s)"), med


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 8))
    base, med = cpu_arm(steps, max(1, min(args.warmup, 2)))
    line = dict(impl="reference", metric="AV clip-pairs/sec (fwd+bwd, ViT-B pretrain)", value=base["value"], unit="clip-pairs/s",
                n_gpus=args.gpus, steps=steps, warmup=max(1, min(args.warmup, 2)), ms_per_step=med * 1e3, higher_is_better=True,
                scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
                config=dict(workload=WORKLOAD, note="CPU arm: bounded sample of the same model at batch 2 per step (BASELINE configs[0])"),
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
def build_trainer(device, distributed):
    import torch
    from deepavfusion_b200.models import AVMAE, DeepAVFusion
    from deepavfusion_b200.util import lr_sched
    from deepavfusion_b200.util.misc import Trainer
    torch.manual_seed(0)
    enc = DeepAVFusion(image_arch="vit_base", image_pretrained="", image_size=(224, 224),
                       audio_arch="vit_base", audio_pretrained="", audio_size=(128, 192),
                       fusion_arch="factorized_mmi", fusion_layers="all", num_fusion_tkns=(16, 8, 8),
                       fusion_mlp_ratio=1.0, fusion_attn_ratio=0.25, fusion_num_heads=12)
    model = AVMAE(enc, enc.embed_dim, image_decoder_depth=8, image_mask_ratio=0.75, image_norm_loss=True,
                  audio_decoder_depth=8, audio_mask_ratio=0.8, audio_norm_loss=True).to(device)
    no_wd = [n for n, p in model.named_parameters() if "bias" in n or "norm" in n]              # train.py:88
    groups = lr_sched.param_groups_pretrained(model, 0.05, no_weight_decay_list=no_wd, image_pt="vit_base_mae_in1k", audio_pt="vit_base_audiomae_as2m")
    lr = 1.5e-4 * BATCH_PER_GPU * max(1, int(os.environ.get("WORLD_SIZE", "1"))) / 256           # train.py:33-35
    opt = torch.optim.AdamW(groups, lr=lr, betas=(0.9, 0.95))                                    # train.py:93
    return Trainer(model, optimizer=opt, accum_iter=1, use_amp=True, distributed=distributed)


def synth_inputs(batch, seed, pinned):
    import torch
    g = torch.Generator().manual_seed(seed)
    image = torch.randn(batch, 3, 224, 224, generator=g)
    audio = (1.5 * torch.randn(batch, 1, 128, 192, generator=g) - 3).clamp_(-7, 3)               # log-mel-like
    if pinned:
        image, audio = image.pin_memory(), audio.pin_memory()
    return image, audio


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
    return dict(bound="tensor", achieved=achieved, peak=peaks["tflops"], unit="TFLOP/s", frac=achieved / peaks["tflops"], traffic=None,
                kernel="davf::gemm_tc_kernel (tcgen05/TMA)", launches_per_step=len(calls), gemm_ms_per_step=sec * 1e3,
                flops_per_launch=flops / len(calls), avg_launch_us=sec * 1e6 / len(calls), peak_source=peaks["src"]), sec


def run_ours(args):
    import torch
    import torch.distributed as dist
    import deepavfusion_b200.kernels as K
    from deepavfusion_b200.util import distributed as dist_utils
    from deepavfusion_b200.util.graphed import GraphedTrainStep

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = dist_utils.init_from_env("nccl") if world > 1 else 0
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    assert args.gpus == world, f"--gpus {args.gpus} but WORLD_SIZE={world} (launch with torchrun for N > 1)"
    peaks = measured_peaks()

    trainer = build_trainer(dev, distributed=world > 1)
    h_image, h_audio = synth_inputs(BATCH_PER_GPU, 1000 + rank, pinned=True)
    d_image, d_audio = h_image.to(dev), h_audio.to(dev)
    torch.manual_seed(2000 + rank)
    use_graph = not args.no_graph

    def eager_step(image, audio):
        li, la, _, _ = trainer.model(image, audio)
        norm, _ = trainer.step(li + la)
        return li, la, norm

    # one traced eager step: GEMM launch list + launch count per step
    K.GEMM_TRACE = []
    n0 = K.launch_count()
    eager_step(d_image, d_audio)
    torch.cuda.synchronize()
    launches_per_step = K.launch_count() - n0
    trace, K.GEMM_TRACE = K.GEMM_TRACE, None

    if args.profile_step:                     # for `ncu --profile-from-start off`: exactly one eager step is profiled
        eager_step(d_image, d_audio)
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
        eager_step(d_image, d_audio)
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
        return

    if use_graph:
        gstep = GraphedTrainStep(trainer, d_image, d_audio, warmup=2)
        step = gstep
        launches_per_step = gstep.launches_per_step
    else:
        step = eager_step

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    # ---- device-resident throughput --------------------------------------------------------------
    last = {}

    def dev_step():
        last["out"] = step(d_image, d_audio)
    with ClockSampler(local) as cs:                 # nvidia-smi needs ~0.3 s to start: launched before the warm-up steps,
        for _ in range(max(3, args.warmup)):        # only rows sampled during the timed region are summarised
            dev_step()
        torch.cuda.synchronize()
        cs.mark()
        ms = timed(dev_step, args.steps)
    clocks = cs.summary()
    li, la, norm = last["out"]
    loss_val = float(li) + float(la)
    assert loss_val == loss_val and abs(loss_val) < 1e4, f"non-finite loss {loss_val}"
    ms_per_step = ms / args.steps
    value = BATCH_PER_GPU * world / (ms_per_step / 1e3)

    # ---- end to end: pinned host inputs -> H2D -> step -> D2H loss, every step ---------------------
    sink = []

    def e2e_step():
        if use_graph:
            # pinned host inputs -> H2D -> graph replay -> async D2H of the losses; the loss of the PREVIOUS step is
            # read on the host while this one runs (every step's loss is read, one step late: train.py:166 logs it)
            step.step_async(h_image, h_audio)
            if step.pending() > 1:
                li_, la_, _ = step.pop_metrics()
                sink.append(li_ + la_)
        else:
            li_, la_, _ = step(h_image.to(dev, non_blocking=True), h_audio.to(dev, non_blocking=True))
            sink.append((li_ + la_).item())             # D2H read of the step's loss (train.py:166)

    def e2e_drain():
        while use_graph and step.pending() > 0:
            li_, la_, _ = step.pop_metrics()
            sink.append(li_ + la_)
    for _ in range(3):
        e2e_step()
    e2e_drain()

    def e2e_run(n):                                     # n steps launched AND their n losses read inside the timed region
        for _ in range(n):
            e2e_step()
        e2e_drain()
    e2e_ms = timed(lambda: e2e_run(args.steps), 1) / args.steps
    assert len(sink) >= args.steps and all(v == v for v in sink[-args.steps:])
    e2e_value = BATCH_PER_GPU * world / (e2e_ms / 1e3)
    h2d = h_image.numel() * 4 + h_audio.numel() * 4

    # ---- roofline of the dominant kernel + CPU baseline (rank 0, N = 1 only) ------------------------
    roof, base = None, None
    if rank == 0:
        roof, gemm_sec = gemm_roofline(trace, peaks)
        roof["share_of_step"] = gemm_sec * 1e3 / ms_per_step
        prof = os.path.join(ROOT, "profiles", "gemm_traffic.json")
        if os.path.exists(prof):
            try:
                roof["traffic"] = json.load(open(prof)).get("dram_bytes_per_launch")
            except Exception:
                pass
    if rank == 0 and world == 1 and not args.skip_cpu:
        base, _ = cpu_arm(steps=5, warmup=1)

    if rank == 0:
        step_tflops = value * GFLOP_PER_PAIR / 1e3
        line = dict(metric="AV clip-pairs/sec (fwd+bwd, ViT-B pretrain)", value=value, unit="clip-pairs/s", n_gpus=world,
                    steps=args.steps, warmup=max(3, args.warmup), ms_per_step=ms_per_step, higher_is_better=True, scaling="weak",
                    vs_baseline=None, dtype="bf16", data="synthetic",
                    config=dict(workload=WORKLOAD, global_batch=BATCH_PER_GPU * world, parallelism=f"dp{world}",
                                step="mask + fwd + bwd + grad all-reduce + fused AdamW", cuda_graph=bool(use_graph),
                                streams=os.environ.get("DAVF_STREAMS", "1") != "0",
                                allreduce=("none" if world == 1 else (("bucketed NCCL all-reduce captured inside the step graph, overlapped with backward" if getattr(step, "overlap_comm", False)
                                                                      else "one NCCL call after the captured fwd+bwd graph") if use_graph else "bucketed, overlapped with backward")),
                                l2="per-step working set (640 MB bf16 weights + >4 GB activations) exceeds the 126 MB L2; no flush needed",
                                gflop_per_pair=GFLOP_PER_PAIR, step_tflops=step_tflops, step_frac_of_peak=step_tflops / (peaks["tflops"] * world)),
                    e2e=dict(value=e2e_value, unit="clip-pairs/s", h2d_bytes_per_step=h2d * world, d2h_bytes_per_step=(12 if use_graph else 4) * world,
                             readback=("pinned async D2H of (loss_image, loss_audio, grad_norm) every step, read on the host one step later" if use_graph else "loss.item() every step"),
                             ms_per_step=e2e_ms),
                    gpu_launches=int(launches_per_step * args.steps), launches_per_step=int(launches_per_step),
                    clocks=clocks, roofline=roof, loss=loss_val)
        if base is not None:
            line["cpu_baseline"] = base
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
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of the whole-step CUDA graph")
    ap.add_argument("--skip-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--profile-step", action="store_true", help="run one eager step between cudaProfilerStart/Stop and exit (for ncu)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Kernel timeline of ONE graph-replayed training step (torch.profiler / CUPTI): per-stream busy time, span,
and the largest gaps.  Run on the GPU box; writes gpurun_out/timeline_summary.txt."""
import collections, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from deepavfusion_b200.util.graphed import GraphedTrainStep
WORLD, RANK = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))     # under torchrun: every rank steps, rank 0 profiles
LOCAL = 0
if WORLD > 1:
    from deepavfusion_b200.util import distributed as dist_utils
    LOCAL = dist_utils.init_from_env("nccl")
dev = torch.device("cuda", LOCAL); torch.cuda.set_device(LOCAL)
CFG = bench.CONFIGS[os.environ.get("DAVF_BENCH_CONFIG", "vggsound")]
trainer = bench.build_trainer(CFG, dev, WORLD > 1)
img, aud = bench.synth_inputs(CFG, CFG["batch"], 1000 + RANK, False)[:2]
img, aud = img.to(dev), aud.to(dev)
def _eager():
    li, la, _, _ = trainer.model(img, aud); trainer.step(li + la)
_eager()   # outputs of eager steps must not outlive the capture (see GraphedTrainStep docstring)
g = GraphedTrainStep(trainer, img, aud, warmup=2)
for _ in range(3): g(img, aud)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
NSTEPS = int(os.environ.get("DAVF_TIMELINE_STEPS", "1"))      # > 1: back-to-back replays, the host running ahead as in training
if RANK != 0:
    for _ in range(NSTEPS): g(img, aud)
    torch.cuda.synchronize(); torch.distributed.barrier(); sys.exit(0)
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(NSTEPS):
        g(img, aud)
    torch.cuda.synchronize()
if WORLD > 1: torch.distributed.barrier()
os.makedirs("gpurun_out", exist_ok=True)
prof.export_chrome_trace("gpurun_out/timeline.json")
ev = json.load(open("gpurun_out/timeline.json"))["traceEvents"]
ks = [e for e in ev if e.get("cat") == "kernel"]
ks.sort(key=lambda e: e["ts"])
if NSTEPS > 1:
    # steady state: the gap between the last kernel of one replay and the first kernel of the next (the graph's head), then
    # keep the LAST replay for the per-stream summary below
    starts = [i for i, e in enumerate(ks) if "distribution_elementwise" in e["name"]][::2]      # first rand kernel of each step
    gaps = []
    for i in starts[1:]:
        prev_end = max(e["ts"] + e["dur"] for e in ks[:i] if "Fill" not in e["name"] or e["ts"] < ks[i]["ts"] - 2000)
        gaps.append((ks[i]["ts"] - prev_end) / 1e3)
    print("steady-state head gaps between consecutive replays (ms):", [round(x, 3) for x in gaps])
    print("step periods (ms):", [round((ks[b]["ts"] - ks[a]["ts"]) / 1e3, 3) for a, b in zip(starts, starts[1:])])
    ks = ks[starts[-1]:]
t0 = ks[0]["ts"]; t1 = max(e["ts"] + e["dur"] for e in ks)
out = [f"kernels {len(ks)}  span {(t1 - t0) / 1e3:.2f} ms"]
by = collections.defaultdict(list)
for e in ks: by[e["args"].get("stream")].append(e)
for s, L in by.items():
    busy = sum(e["dur"] for e in L)
    out.append(f"stream {s}: {len(L)} kernels, busy {busy / 1e3:.2f} ms, first {(L[0]['ts'] - t0) / 1e3:.2f} last {(L[-1]['ts'] + L[-1]['dur'] - t0) / 1e3:.2f}")
# machine occupancy over time: how much of the span has >=1 kernel running, and time with only 'small' kernels
events = []
for e in ks: events += [(e["ts"], 1), (e["ts"] + e["dur"], -1)]
events.sort()
cur = 0; last = t0; idle = 0.0
for t, d in events:
    if cur == 0: idle += t - last
    cur += d; last = t
out.append(f"time with no kernel running: {idle / 1e3:.2f} ms")
# per-phase: find the adamw kernel and the first backward kernel (masked_mse bwd)
def first(name):
    for e in ks:
        if name in e["name"]: return (e["ts"] - t0) / 1e3
out.append(f"loss bwd starts at {first('masked_mse_kernel<true>')} ms, adamw at {first('adamw_kernel')} ms")
agg = collections.defaultdict(float)
for e in ks: agg[e["name"][:60]] += e["dur"]
for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:24]: out.append(f"{v / 1e3:8.2f} ms  {k}")
# the head of the step: the first kernels and the gaps between them
out.append("first 40 kernels (start ms, dur us, stream, name):")
for e in ks[:40]: out.append(f"  {(e['ts'] - t0) / 1e3:7.3f} {e['dur']:7.1f} {e['args'].get('stream')} {e['name'][:70]}")
out.append(f"sum of kernel time {sum(e['dur'] for e in ks) / 1e3:.2f} ms")
if WORLD > 1:
    # the exchange step: every NCCL kernel and the AdamW launch that follows it, against the end of backward
    out.append("exchange step (start ms, dur us, name):")
    for e in ks:
        if "nccl" in e["name"].lower() or "adamw" in e["name"]:
            out.append(f"  {(e['ts'] - t0) / 1e3:7.3f} {e['dur']:8.1f} {e['name'][:60]}")
    nc = [e for e in ks if "nccl" in e["name"].lower()]
    other = [e for e in ks if "nccl" not in e["name"].lower() and "adamw" not in e["name"]]
    out.append(f"NCCL busy {sum(e['dur'] for e in nc) / 1e3:.2f} ms over {len(nc)} kernels; last compute kernel ends at "
               f"{max(e['ts'] + e['dur'] for e in other) / 1e3 - t0 / 1e3:.2f} ms, step ends at {(t1 - t0) / 1e3:.2f} ms")
SUFFIX = os.environ.get("DAVF_TIMELINE_TAG", "")
open(f"gpurun_out/timeline_summary{SUFFIX}.txt", "w").write("\n".join(out))
print("\n".join(out))

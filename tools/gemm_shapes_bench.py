#!/usr/bin/env python
"""Per-shape timing of the step's GEMM launches (run on the GPU box): traces one eager step of the
bench workload, groups launches by signature and times each signature hot (CUDA events, 20 reps)."""
import collections, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import deepavfusion_b200.kernels as K

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
CFG = bench.CONFIGS[os.environ.get("DAVF_BENCH_CONFIG", "vggsound")]
trainer = bench.build_trainer(CFG, dev, False)
img, aud = bench.synth_inputs(CFG, CFG["batch"], 1000, False)[:2]
img, aud = img.to(dev), aud.to(dev)
K.GEMM_TRACE = []
li, la, _, _ = trainer.model(img, aud)
trainer.step(li + la)
torch.cuda.synchronize()
trace, K.GEMM_TRACE = K.GEMM_TRACE, None
groups = collections.OrderedDict()
for t in trace:
    key = json.dumps({k: t[k] for k in sorted(t)}, sort_keys=True, default=str)
    groups.setdefault(key, [t, 0])[1] += 1
rows = []
peaks = bench.measured_peaks()
for key, (t, n) in groups.items():
    roof, sec = bench.gemm_roofline([t], peaks, reps=20)
    rows.append((sec * n, sec, n, t))
rows.sort(key=lambda r: -r[0])
tot = sum(r[0] for r in rows)
print(f"total GEMM time/step {tot*1e3:.2f} ms over {sum(r[2] for r in rows)} launches, {len(rows)} signatures")
for tt, sec, n, t in rows[:70]:
    if "group" in t:
        fl = sum(2.0 * u["M"] * u["N"] * u["K"] for u in t["group"])
        u0 = t["group"][0]
        maj = ("K" if u0["a_kmajor"] else "MN") + "/" + ("K" if u0["b_kmajor"] else "MN")
        print(f"{tt*1e3:7.3f} ms  n={n:3d}  {sec*1e6:7.1f} us  {fl/sec/1e12:7.1f} TF/s  GROUP {maj} " + " ".join(f"{u['M']}x{u['N']}x{u['K']}" for u in t["group"]))
        continue
    fl = 2.0 * t["M"] * t["N"] * t["K"]
    maj = ("K" if t["a_kmajor"] else "MN") + "/" + ("K" if t["b_kmajor"] else "MN")
    flags = "".join(c for c, on in (("b", t["bias"]), ("G", t["act"] == 1), ("D", t["act"] == 2), ("x", t["aux"]), ("r", t["res"]), ("A", t["accumulate"]), ("s", t.get("rowsum"))) if on)
    print(f"{tt*1e3:7.3f} ms  n={n:3d}  {sec*1e6:7.1f} us  {fl/sec/1e12:7.1f} TF/s  M={t['M']:6d} N={t['N']:5d} K={t['K']:6d} {maj:6s} {'bf16' if t['out_bf16'] else 'f32 '} {flags}")

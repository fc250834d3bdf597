#!/usr/bin/env python
"""Launch a handful of representative GEMM signatures (for `ncu --set full -k regex:gemm_tc`)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import deepavfusion_b200.kernels as K
bf16 = torch.bfloat16
def r(*s, dt=bf16): return (torch.randn(*s, device="cuda") * 0.05).to(dt)
cases = {
    "plain_big":  lambda: K.gemm(r(14592, 2048), r(2048, 512), True, False),                       # dgrad fc2-like, best case so far
    "bias_qkv":   lambda: K.gemm(r(14592, 512), r(1536, 512), bias=r(1536, dt=torch.float32)),     # fwd qkv decoder
    "gelu_fc1":   lambda: K.gemm(r(14592, 512), r(2048, 512), bias=r(2048, dt=torch.float32), act=K.ACT_GELU, want_aux=True),
    "res_proj":   lambda: K.gemm(r(14592, 512), r(512, 512), bias=r(512, dt=torch.float32), res=r(14592, 512, dt=torch.float32), out_dtype=torch.float32),
    "wgrad":      lambda: K.gemm(r(3136, 768), r(3136, 768), False, False, out=torch.zeros(768, 768, device="cuda"), accumulate=True),
    "small":      lambda: K.gemm(r(512, 768), r(768, 768), bias=r(768, dt=torch.float32)),
}
which = [a for a in sys.argv[1:] if a in cases] or ([] if 'timeline' in sys.argv else list(cases))
for name in which:
    f = cases[name]
    for _ in range(int(os.environ.get("PROBE_REPS", "3"))):
        f()
    torch.cuda.synchronize()
    print("ran", name)

# timeline probe of CTA 0 (SM clocks relative to kernel start)
if "timeline" in sys.argv or len(sys.argv) == 1:
    for (M, N, Kd, tag) in [(512, 768, 768, "small"), (14592, 1536, 512, "qkv_dec"), (3136, 768, 768, "enc")]:
        a, b, bias = r(M, Kd), r(N, Kd), r(N, dt=torch.float32)
        dbg = torch.zeros(8, dtype=torch.int64, device="cuda")
        for _ in range(3):
            K.gemm(a, b, bias=bias, debug_clocks=dbg)
        torch.cuda.synchronize()
        d = dbg.cpu().tolist()
        names = ["start", "producer_first_issue", "mma_first_data", "mma_tile0_issued", "epi_tile0_acc_ready", "epi_tile0_done", "all_done"]
        print(tag, {n: d[i] - d[0] for i, n in enumerate(names)})

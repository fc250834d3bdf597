#!/usr/bin/env python
"""One representative wgrad launch (image-block qkv: dW[2304,768] += dy^T x, K = 5184 rows, + bias-gradient row sums)
for `ncu --set full -k regex:gemm_tc -s 2 -c 1 python tools/wgrad_probe.py`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import deepavfusion_b200.kernels as K
bf16 = torch.bfloat16
def r(*s, dt=bf16): return (torch.randn(*s, device="cuda") * 0.05).to(dt)
# wgrad qkv image: dW[2304,768] += dy[5184,2304]^T x[5184,768]  (+ bias grad)
dy, x = r(5184, 2304), r(5184, 768)
out = torch.zeros(2304, 768, device="cuda"); db = torch.zeros(2304, device="cuda")
for _ in range(3):
    K.gemm(dy, x, False, False, out=out, accumulate=True, rowsum_out=db)
torch.cuda.synchronize()
print("done")

#!/usr/bin/env python
"""The step's top time consumer, the grouped CTA-pair wgrad launch gemm_tc_kernel<256,4,0,0,192,2,8,6> (a ViT block's two
weight gradients + bias-gradient row sums in one launch), and one fusion-block group, for
  ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 6 -c 3 -o gpurun_out/prof_wgrad_group python tools/wgrad_group_probe.py
Launches after the warm-up: image-block MLP wgrads (fc2 + fc1), decoder-image MLP wgrads, fusion-block forward group."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import deepavfusion_b200.kernels as K
bf16 = torch.bfloat16
def r(*s, dt=bf16): return (torch.randn(*s, device="cuda") * 0.05).to(dt)


def wgrad_pair(rows, d, hidden):
    dy2, a = r(rows, d), r(rows, hidden)          # fc2: dW[d, hidden] += dy^T a
    dh, xn = r(rows, hidden), r(rows, d)          # fc1: dW[hidden, d] += dh^T xn
    w2, b2 = torch.zeros(d, hidden, device="cuda"), torch.zeros(d, device="cuda")
    w1, b1 = torch.zeros(hidden, d, device="cuda"), torch.zeros(hidden, device="cuda")
    return [((dy2, a, False, False), dict(out=w2, accumulate=True, rowsum_out=b2)), ((dh, xn, False, False), dict(out=w1, accumulate=True, rowsum_out=b1))]


def fusion_group():
    mv, xv, ma, xa, m2 = r(512, 768), r(3136, 768), r(512, 768), r(1216, 768), r(1024, 768)
    ws = [r(768, 768), r(1536, 768), r(768, 768), r(1536, 768), r(192, 768)]
    bs = [torch.zeros(w.shape[0], device="cuda") for w in ws]
    return [((x, w, True, True), dict(bias=b)) for x, w, b in zip((mv, xv, ma, xa, m2), ws, bs)]


groups = [wgrad_pair(3136, 768, 3072), wgrad_pair(14592, 512, 2048), fusion_group()]
for _ in range(3):
    for g in groups:
        K.gemm_grouped(g)
torch.cuda.synchronize()
print("done")

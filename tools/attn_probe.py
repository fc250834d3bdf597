#!/usr/bin/env python
"""The tcgen05 attention kernels at the bench workload's shapes, for
  ncu --set full --clock-control none --import-source on -k regex:attn_tc -s 8 -c 8 -o gpurun_out/prof_attn python tools/attn_probe.py
Launch order after the 8 warm-up launches: decoder image fwd, bwd; decoder audio fwd, bwd; encoder image fwd, bwd; encoder audio fwd, bwd."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import deepavfusion_b200.kernels as K
bf16 = torch.bfloat16


def run(B, H, S, nP, hd):
    n = S - nP
    qkv = torch.randn(B, S, 3, H, hd, device="cuda").to(bf16)
    dqkv = torch.empty_like(qkv)
    do = torch.randn(B, n, H, hd, device="cuda").to(bf16)
    o, lse = K.attention_fwd(qkv[:, nP:, 0], qkv[:, :, 1], qkv[:, :, 2], hd ** -0.5)
    K.attention_bwd(qkv[:, nP:, 0], qkv[:, :, 1], qkv[:, :, 2], do, lse, hd ** -0.5, dqkv[:, nP:, 0], dqkv[:, :, 1], dqkv[:, :, 2], o=o, dq_dead_rows=nP)


for _ in range(2):
    run(64, 16, 228, 0, 32)
    run(64, 16, 128, 0, 32)
    run(64, 12, 81, 32, 64)
    run(64, 12, 51, 32, 64)
torch.cuda.synchronize()
print("done")

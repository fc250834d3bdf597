#!/usr/bin/env python
"""Hot timings (CUDA events, L2-cold by rotating buffers) of the non-GEMM kernels at the bench workload's shapes:
attention fwd / bwd, LayerNorm fwd / bwd.  Run on the GPU box."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import deepavfusion_b200.kernels as K

bf16 = torch.bfloat16
dev = "cuda"
NBUF = 6      # rotate over buffers so that inputs are not L2-resident between iterations


def timeit(fn, reps=5):
    """fn(i) launches on the current stream; NBUF * 4 launches are captured into a CUDA graph (the Python / ctypes
    launch path costs ~30 us per call, more than most of these kernels) and the graph is replayed."""
    for i in range(NBUF):
        fn(i)
    torch.cuda.synchronize()
    n = NBUF * 4
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(n):
            fn(i)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (reps * n) * 1e3


def attn(B, H, S, nP, hd, tag):
    n = S - nP
    qkvs = [torch.randn(B, S, 3, H, hd, device=dev).to(bf16) for _ in range(NBUF)]
    dos = [torch.randn(B, n, H, hd, device=dev).to(bf16) for _ in range(NBUF)]
    dqkv = [torch.empty_like(q) for q in qkvs]
    outs = [K.attention_fwd(q[:, nP:, 0], q[:, :, 1], q[:, :, 2], hd ** -0.5) for q in qkvs]
    t_f = timeit(lambda i: K.attention_fwd(qkvs[i % NBUF][:, nP:, 0], qkvs[i % NBUF][:, :, 1], qkvs[i % NBUF][:, :, 2], hd ** -0.5))

    def bwd(i):
        j = i % NBUF
        q5, d5 = qkvs[j], dqkv[j]
        K.attention_bwd(q5[:, nP:, 0], q5[:, :, 1], q5[:, :, 2], dos[j], outs[j][1], hd ** -0.5,
                        d5[:, nP:, 0], d5[:, :, 1], d5[:, :, 2], o=outs[j][0])
    t_b = timeit(bwd)
    fl = 2.0 * B * H * n * S * 2 * hd
    byt_f = (B * S * 2 * H * hd + 2 * B * n * H * hd) * 2
    byt_b = (B * S * 3 * H * hd * 2 + B * n * H * hd * 2) * 2
    print(f"attn {tag:14s} B{B} H{H} Nq{n} Nk{S} d{hd}: fwd {t_f:7.1f} us ({fl / t_f / 1e6:6.1f} TF/s, {byt_f / t_f / 1e3:6.0f} GB/s)   "
          f"bwd {t_b:7.1f} us ({2.5 * fl / t_b / 1e6:6.1f} TF/s, {byt_b / t_b / 1e3:6.0f} GB/s)")


def ln(B, n0, n1, D, tag):
    x0 = [torch.randn(B, n0, D, device=dev) for _ in range(NBUF)]
    x1 = [torch.randn(B, n1, D, device=dev) for _ in range(NBUF)] if n1 else [None] * NBUF
    g, b = torch.randn(D, device=dev), torch.randn(D, device=dev)
    rows = B * (n0 + n1)
    outs = [K.layernorm_fwd(x0[j], x1[j], g, b, 1e-6) for j in range(NBUF)]
    t_f = timeit(lambda i: K.layernorm_fwd(x0[i % NBUF], x1[i % NBUF], g, b, 1e-6))
    dyb = [torch.randn(rows, D, device=dev).to(bf16) for _ in range(NBUF)]
    add0 = [torch.randn(B, n0, D, device=dev) for _ in range(NBUF)]
    dg, db = torch.zeros(D, device=dev), torch.zeros(D, device=dev)

    def bwd(i):
        j = i % NBUF
        K.layernorm_bwd(x0[j], x1[j], g, outs[j][2], outs[j][3], dyb[j], None, add0[j], None, dg, db)
    t_b = timeit(bwd)
    print(f"ln   {tag:14s} rows {rows} D{D}: fwd {t_f:6.1f} us ({rows * D * 6 / t_f / 1e3:6.0f} GB/s)   bwd {t_b:6.1f} us ({rows * D * 14 / t_f / 1e3 * t_f / t_b:6.0f} GB/s)")


if __name__ == "__main__":
    torch.cuda.set_device(0)
    for impl, name in ((0, "tcgen05 (product dispatch)"), (2, "mma.sync everywhere")):
        K.set_attn_impl(impl)
        print(f"--- attention impl {impl}: {name}")
        attn(64, 12, 81, 32, 64, "enc image")
        attn(64, 12, 51, 32, 64, "enc audio")
        attn(64, 16, 228, 0, 32, "dec image")
        attn(64, 16, 128, 0, 32, "dec audio")
        attn(32, 12, 228, 32, 64, "enc image full")
        attn(32, 12, 128, 32, 64, "enc audio full")
        attn(64, 12, 49, 41, 64, "cross v (8q)")
        attn(64, 12, 19, 11, 64, "cross a (8q)")
    K.set_attn_impl(0)
    ln(64, 32, 49, 768, "enc image")
    ln(64, 49, 0, 768, "enc image mlp")
    ln(64, 228, 0, 512, "dec image")
    ln(64, 128, 0, 512, "dec audio")

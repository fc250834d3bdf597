"""Tensor-level wrappers over the C ABI (include/davf.h).

PyTorch is used here for device memory and streams only: every wrapper checks dtypes / layouts,
allocates outputs from the torch caching allocator, and launches the hand-written sm_100a kernels
of libdavf_sm100.so on the current torch CUDA stream with raw device pointers.  There is no
fallback: a missing library or a CPU tensor raises.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import torch

from . import _cabi
from ._cabi import check

Tensor = torch.Tensor
ACT_NONE, ACT_GELU, ACT_DGELU = 0, 1, 2


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t: Optional[Tensor]) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


def _need(t: Tensor, dtype, name: str, contiguous: bool = True) -> Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"{name}: expected a CUDA tensor (the sm_100a kernels have no CPU path)")
    if t.dtype != dtype:
        raise TypeError(f"{name}: expected {dtype}, got {t.dtype}")
    if contiguous and not t.is_contiguous():
        raise ValueError(f"{name}: expected a contiguous tensor, got strides {t.stride()}")
    return t


# --------------------------------------------------------------------------------------------
# K2 masking
# --------------------------------------------------------------------------------------------
def mask_rank(noise: Tensor, len_keep: int) -> Tuple[Tensor, Tensor, Tensor]:
    """noise f32 [B,L] -> (ids_restore i64 [B,L], ids_keep i64 [B,len_keep], mask f32 [B,L])."""
    _need(noise, torch.float32, "noise")
    B, L = noise.shape
    ids_restore = torch.empty(B, L, dtype=torch.int64, device=noise.device)
    ids_keep = torch.empty(B, len_keep, dtype=torch.int64, device=noise.device)
    mask = torch.empty(B, L, dtype=torch.float32, device=noise.device)
    check(_cabi.lib().davf_mask_rank(_ptr(noise), B, L, len_keep, _ptr(ids_restore), _ptr(ids_keep), _ptr(mask), _stream()), "davf_mask_rank")
    return ids_restore, ids_keep, mask


# --------------------------------------------------------------------------------------------
# row kernels
# --------------------------------------------------------------------------------------------
def patch_rows(img: Tensor, ids_keep: Optional[Tensor], p: int) -> Tensor:
    """img f32 [B,C,H,W] -> bf16 [B*nK, C*p*p] im2col rows (c,py,px order) of the kept patches."""
    _need(img, torch.float32, "img")
    B, Cc, H, W = img.shape
    nK = (H // p) * (W // p) if ids_keep is None else ids_keep.shape[1]
    if ids_keep is not None:
        _need(ids_keep, torch.int64, "ids_keep")
    out = torch.empty(B * nK, Cc * p * p, dtype=torch.bfloat16, device=img.device)
    check(_cabi.lib().davf_patch_rows(_ptr(img), _ptr(ids_keep), _ptr(out), B, Cc, H, W, p, nK, _stream()), "davf_patch_rows")
    return out


def cast_rows_bf16(src: Tensor, M: Optional[int] = None, g: Optional[int] = None, G: Optional[int] = None, off: int = 0) -> Tensor:
    """f32 rows -> bf16 [M, D]; output row m reads source row (m//g)*G + off + m%g."""
    _need(src, torch.float32, "src")
    D = src.shape[-1]
    rows = src.numel() // D
    if M is None:
        M, g, G, off = rows, max(rows, 1), max(rows, 1), 0
    out = torch.empty(M, D, dtype=torch.bfloat16, device=src.device)
    check(_cabi.lib().davf_cast_rows_bf16(_ptr(src), _ptr(out), M, D, g, G, off, _stream()), "davf_cast_rows_bf16")
    return out


def cast_flat_bf16(src: Tensor, dst: Tensor) -> None:
    _need(src, torch.float32, "src"); _need(dst, torch.bfloat16, "dst")
    check(_cabi.lib().davf_cast_flat_bf16(_ptr(src), _ptr(dst), src.numel(), _stream()), "davf_cast_flat_bf16")


def sum_cast(parts: Sequence[Tensor], want_bf16: bool = True):
    """f32 sum of two or three equally shaped tensors, plus its bf16 copy: (sum, bf16 or None)."""
    assert 2 <= len(parts) <= 3
    a = _need(parts[0], torch.float32, "parts[0]")
    for t in parts[1:]:
        _need(t, torch.float32, "parts[i]")
        assert t.shape == a.shape
    out = torch.empty_like(a)
    lp = torch.empty(a.shape, dtype=torch.bfloat16, device=a.device) if want_bf16 else None
    check(_cabi.lib().davf_sum_cast(_ptr(a), _ptr(parts[1]), _ptr(parts[2] if len(parts) > 2 else None), _ptr(out), _ptr(lp), a.numel(), _stream()),
          "davf_sum_cast")
    return out, lp


def colsum_bf16(x: Tensor, out: Tensor) -> None:
    """out[n] += sum_m x[m, n]   (x bf16 [M,N] row-major view with stride(1) == 1)."""
    _need(x, torch.bfloat16, "x", contiguous=False); _need(out, torch.float32, "out")
    assert x.dim() == 2 and x.stride(1) == 1
    check(_cabi.lib().davf_colsum_bf16(_ptr(x), x.shape[0], x.shape[1], x.stride(0), _ptr(out), _stream()), "davf_colsum_bf16")


def batchsum_f32(x: Tensor, off: int, g: int, out: Tensor, accumulate: bool) -> None:
    """out[r, :] (+)= sum_b x[b, off + r, :]   (x f32 [B,G,D], out f32 [g,D])."""
    _need(x, torch.float32, "x"); _need(out, torch.float32, "out")
    B, G, D = x.shape
    check(_cabi.lib().davf_batchsum_f32(_ptr(x), B, G, off, g, D, _ptr(out), int(accumulate), _stream()), "davf_batchsum_f32")


# --------------------------------------------------------------------------------------------
# K4 LayerNorm
# --------------------------------------------------------------------------------------------
def _bstride(x: Tensor) -> int:
    """batch stride of a [B,n,D] tensor whose rows are dense (stride 0 = broadcast sample)."""
    assert x.dim() == 3 and x.stride(2) == 1 and (x.shape[1] == 1 or x.stride(1) == x.shape[2]), x.stride()
    return x.stride(0) if x.shape[0] > 1 else x.shape[1] * x.shape[2]


def _segs(arr, seg_start):
    n = 0
    if seg_start is not None and len(seg_start) > 2:
        n = len(seg_start) - 1
        for k, v in enumerate(seg_start):
            arr[k] = int(v)
    return n


def layernorm_fwd(x0: Tensor, x1: Optional[Tensor], gamma: Tensor, beta: Tensor, eps: float,
                  want_bf16: bool = True, want_f32: bool = False, seg_start: Optional[Sequence[int]] = None):
    """LayerNorm over the rows of cat([x0, x1], dim=1).  Returns (y_bf16, y_f32, mean, rstd);
    y_* are [B*(n0+n1), D] (bf16 rows are segment-major when seg_start has more than one segment)."""
    _need(x0, torch.float32, "x0", contiguous=False)
    B, n0, D = x0.shape
    n1 = 0
    if x1 is not None:
        _need(x1, torch.float32, "x1", contiguous=False)
        assert x1.shape[0] == B and x1.shape[2] == D
        n1 = x1.shape[1]
    rows = B * (n0 + n1)
    dev = x0.device
    y_bf16 = torch.empty(rows, D, dtype=torch.bfloat16, device=dev) if want_bf16 else None
    y_f32 = torch.empty(rows, D, dtype=torch.float32, device=dev) if want_f32 else None
    mean = torch.empty(rows, dtype=torch.float32, device=dev)
    rstd = torch.empty(rows, dtype=torch.float32, device=dev)
    a = _cabi.LnFwdArgs()
    a.x0, a.bs0, a.n0 = x0.data_ptr(), _bstride(x0), n0
    a.x1, a.bs1, a.n1 = (x1.data_ptr() if x1 is not None else 0), (_bstride(x1) if x1 is not None else 0), n1
    a.B, a.D, a.eps = B, D, float(eps)
    a.gamma, a.beta = _need(gamma, torch.float32, "gamma").data_ptr(), _need(beta, torch.float32, "beta").data_ptr()
    a.y_bf16 = y_bf16.data_ptr() if want_bf16 else 0
    a.y_f32 = y_f32.data_ptr() if want_f32 else 0
    a.mean, a.rstd = mean.data_ptr(), rstd.data_ptr()
    a.nseg = _segs(a.seg_start, seg_start)
    check(_cabi.lib().davf_layernorm_fwd(C.byref(a), _stream()), "davf_layernorm_fwd")
    return y_bf16, y_f32, mean, rstd


def layernorm_bwd(x0: Tensor, x1: Optional[Tensor], gamma: Tensor, mean: Tensor, rstd: Tensor,
                  dy_bf16: Optional[Tensor], dy_f32: Optional[Tensor],
                  add0: Optional[Tensor], add1: Optional[Tensor], dgamma: Tensor, dbeta: Tensor,
                  seg_start: Optional[Sequence[int]] = None, need_dx0: bool = True, need_dx1: bool = True,
                  dx0_out: Optional[Tensor] = None, dx0_lowp: Optional[Tensor] = None, dx1_lowp: Optional[Tensor] = None):
    """Returns (dx0 [B,n0,D] f32, dx1 [B,n1,D] f32 or None); dgamma / dbeta are accumulated in place.
    ``dx0_out`` may be a [B,n0,D] view with a batch stride (rows dense) to write dx0 in place.
    ``dx0_lowp`` / ``dx1_lowp``: optional contiguous bf16 [B*n0, D] / [B*n1, D] tensors that receive a bf16 copy of
    the dx rows (the GEMM operand of the next backward region)."""
    B, n0, D = x0.shape
    n1 = x1.shape[1] if x1 is not None else 0
    dev = x0.device
    if dx0_out is not None:
        _need(dx0_out, torch.float32, "dx0_out", contiguous=False)
        assert add0 is None or dx0_out.is_contiguous()
        dx0 = dx0_out
    else:
        dx0 = torch.empty(B, n0, D, dtype=torch.float32, device=dev) if need_dx0 else None
    dx1 = torch.empty(B, n1, D, dtype=torch.float32, device=dev) if (x1 is not None and need_dx1) else None
    a = _cabi.LnBwdArgs()
    a.x0, a.bs0, a.n0 = x0.data_ptr(), _bstride(x0), n0
    a.x1, a.bs1, a.n1 = (x1.data_ptr() if x1 is not None else 0), (_bstride(x1) if x1 is not None else 0), n1
    a.B, a.D = B, D
    a.gamma, a.mean, a.rstd = gamma.data_ptr(), mean.data_ptr(), rstd.data_ptr()
    a.dy_bf16 = _need(dy_bf16, torch.bfloat16, "dy_bf16").data_ptr() if dy_bf16 is not None else 0
    a.dy_f32 = _need(dy_f32, torch.float32, "dy_f32").data_ptr() if dy_f32 is not None else 0
    a.dx0, a.dbs0 = (dx0.data_ptr() if dx0 is not None else 0), (_bstride(dx0) if dx0 is not None else n0 * D)
    a.add0 = _need(add0, torch.float32, "add0").data_ptr() if add0 is not None else 0
    a.dx1, a.dbs1 = (dx1.data_ptr() if dx1 is not None else 0), n1 * D
    a.add1 = _need(add1, torch.float32, "add1").data_ptr() if add1 is not None else 0
    a.dgamma, a.dbeta = _need(dgamma, torch.float32, "dgamma").data_ptr(), _need(dbeta, torch.float32, "dbeta").data_ptr()
    a.nseg = _segs(a.seg_start, seg_start)
    if dx0_lowp is not None:
        assert dx0 is not None and dx0_lowp.numel() == B * n0 * D
        a.dx0_bf16 = _need(dx0_lowp, torch.bfloat16, "dx0_lowp").data_ptr()
    if dx1_lowp is not None:
        assert dx1 is not None and dx1_lowp.numel() == B * n1 * D
        a.dx1_bf16 = _need(dx1_lowp, torch.bfloat16, "dx1_lowp").data_ptr()
    check(_cabi.lib().davf_layernorm_bwd(C.byref(a), _stream()), "davf_layernorm_bwd")
    return dx0, dx1


# --------------------------------------------------------------------------------------------
# GEMM
# --------------------------------------------------------------------------------------------
GEMM_TRACE = None     # bench.py sets this to a list to record the (M, N, K, majors, epilogue) of every launch


def _gemm_args(a: Tensor, b: Tensor, a_kmajor: bool = True, b_kmajor: bool = True, *,
               bias: Optional[Tensor] = None, act: int = ACT_NONE, want_aux: bool = False, aux_in: Optional[Tensor] = None,
               res: Optional[Tensor] = None, res_idx: Optional[Tensor] = None,
               out: Optional[Tensor] = None, out_dtype=torch.bfloat16, accumulate: bool = False,
               window: Optional[Tuple[int, int, int]] = None, split_k: int = 0, rowsum_out: Optional[Tensor] = None,
               debug_clocks: Optional[Tensor] = None):
    """Checks one GEMM problem and fills its C-ABI argument block.  Returns (args, result, trace record)."""
    _need(a, torch.bfloat16, "a", contiguous=False); _need(b, torch.bfloat16, "b", contiguous=False)
    assert a.dim() == 2 and b.dim() == 2 and a.stride(1) == 1 and b.stride(1) == 1
    if a_kmajor:
        M, K = a.shape
    else:
        K, M = a.shape
    if b_kmajor:
        N, Kb = b.shape
    else:
        Kb, N = b.shape
    assert K == Kb, f"gemm: reduction mismatch {K} vs {Kb}"
    dev = a.device
    if out is None:
        assert window is None and not accumulate
        out = torch.empty(M, N, dtype=out_dtype, device=dev)
    assert out.stride(-1) == 1
    ga = _cabi.GemmArgs()
    ga.a, ga.lda, ga.a_kmajor = a.data_ptr(), a.stride(0), int(a_kmajor)
    ga.b, ga.ldb, ga.b_kmajor = b.data_ptr(), b.stride(0), int(b_kmajor)
    ga.M, ga.N, ga.K = M, N, K
    ga.bias = _need(bias, torch.float32, "bias").data_ptr() if bias is not None else 0
    ga.act = act
    aux = None
    if want_aux:
        aux = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
        ga.aux_out, ga.ldaux = aux.data_ptr(), N
    if aux_in is not None:
        _need(aux_in, torch.bfloat16, "aux_in")
        ga.aux_in, ga.ldaux = aux_in.data_ptr(), aux_in.shape[-1]
    if res is not None:
        _need(res, torch.float32, "res", contiguous=False)
        assert res.stride(-1) == 1
        ga.res, ga.ldres = res.data_ptr(), res.stride(-2)
    if res_idx is not None:
        ga.res_idx = _need(res_idx, torch.int64, "res_idx").data_ptr()
    ga.out, ga.ldo = out.data_ptr(), out.stride(-2)
    ga.out_bf16 = int(out.dtype == torch.bfloat16)
    assert out.dtype in (torch.bfloat16, torch.float32)
    ga.accumulate = int(accumulate)
    if window is not None:
        ga.g, ga.G, ga.off = window
    ga.split_k = split_k
    if rowsum_out is not None:
        ga.rowsum_out = _need(rowsum_out, torch.float32, "rowsum_out").data_ptr()
        assert rowsum_out.numel() == M
    if debug_clocks is not None:
        ga.debug_clocks = _need(debug_clocks, torch.int64, "debug_clocks").data_ptr()
    rec = None
    if GEMM_TRACE is not None:
        rec = dict(M=M, N=N, K=K, a_kmajor=a_kmajor, b_kmajor=b_kmajor, lda=a.stride(0), ldb=b.stride(0),
                   bias=bias is not None, act=act, aux=want_aux, res=res is not None, res_idx=res_idx is not None,
                   out_bf16=out.dtype == torch.bfloat16, accumulate=accumulate, window=window, ldo=out.stride(-2),
                   rowsum=rowsum_out is not None)
    return ga, ((out, aux) if want_aux else out), rec


def gemm(a: Tensor, b: Tensor, a_kmajor: bool = True, b_kmajor: bool = True, **kw):
    """acc[m,n] = sum_k A(m,k) B(n,k) with the fused epilogue of davf.h.

    ``a`` / ``b`` are the STORED 2-D bf16 matrices (row-major views, stride(1) == 1):
    K-major operand: stored [rows, K]; MN-major operand: stored [K, rows].
    ``window`` = (g, G, off): output (and residual) row of m is (m//g)*G + off + m%g; ``out`` must
    then be given.  Returns ``out`` or ``(out, aux)`` when ``want_aux``.  Keyword options: see ``_gemm_args``."""
    ga, result, rec = _gemm_args(a, b, a_kmajor, b_kmajor, **kw)
    if rec is not None:
        GEMM_TRACE.append(rec)
    check(_cabi.lib().davf_gemm(C.byref(ga), _stream()), "davf_gemm")
    return result


GEMM_MAX_GROUP = 6


def gemm_grouped(problems: Sequence[Tuple[tuple, dict]]):
    """``problems`` = [((a, b, a_kmajor, b_kmajor), kwargs), ...]: independent GEMMs of ONE operand-layout class
    launched as one kernel (davf_gemm_grouped); longer lists are cut into launches of GEMM_MAX_GROUP.  Returns the
    list of per-problem results (as ``gemm`` would)."""
    results = []
    for i0 in range(0, len(problems), GEMM_MAX_GROUP):
        chunk = problems[i0:i0 + GEMM_MAX_GROUP]
        arr = (_cabi.GemmArgs * len(chunk))()
        recs = []
        for j, (pos, kw) in enumerate(chunk):
            ga, result, rec = _gemm_args(*pos, **kw)
            arr[j] = ga
            results.append(result)
            recs.append(rec)
        if GEMM_TRACE is not None:
            GEMM_TRACE.append(dict(group=recs))
        check(_cabi.lib().davf_gemm_grouped(arr, len(chunk), _stream()), "davf_gemm_grouped")
    return results


# --------------------------------------------------------------------------------------------
# attention
# --------------------------------------------------------------------------------------------
def _bhs(t: Tensor, name: str):
    """[B, N, H, d] bf16 view with unit stride on d and head stride d -> (ptr, batch stride, row stride)."""
    _need(t, torch.bfloat16, name, contiguous=False)
    assert t.dim() == 4 and t.stride(3) == 1 and (t.shape[2] == 1 or t.stride(2) == t.shape[3]), (name, t.shape, t.stride())
    return t.data_ptr(), t.stride(0), t.stride(1)


def attention_fwd(q: Tensor, k: Tensor, v: Tensor, scale: float, out: Optional[Tensor] = None, accumulate: bool = False):
    """q [B,Nq,H,dqk], k [B,Nk,H,dqk], v [B,Nk,H,dv] strided bf16 views -> (o [B,Nq,H,dv] bf16, lse f32 [B,H,Nq])."""
    B, Nq, H, dqk = q.shape
    Nk, dv = k.shape[1], v.shape[3]
    if out is None:
        out = torch.empty(B, Nq, H, dv, dtype=torch.bfloat16, device=q.device)
    lse = torch.empty(B, H, Nq, dtype=torch.float32, device=q.device)
    a = _cabi.AttnFwdArgs()
    a.q, a.q_bs, a.q_rs = _bhs(q, "q")
    a.k, a.k_bs, a.k_rs = _bhs(k, "k")
    a.v, a.v_bs, a.v_rs = _bhs(v, "v")
    a.o, a.o_bs, a.o_rs = _bhs(out, "o")
    a.lse = lse.data_ptr()
    a.B, a.H, a.Nq, a.Nk, a.dqk, a.dv = B, H, Nq, Nk, dqk, dv
    a.scale, a.accumulate = float(scale), int(accumulate)
    check(_cabi.lib().davf_attention_fwd(C.byref(a), _stream()), "davf_attention_fwd")
    return out, lse


def attention_bwd(q: Tensor, k: Tensor, v: Tensor, d_o: Tensor, lse: Tensor, scale: float,
                  dq: Tensor, dk: Tensor, dv: Tensor, accumulate_dq: bool = False, o: Optional[Tensor] = None,
                  dq_dead_rows: int = 0) -> None:
    """Writes dq / dk / dv (strided bf16 views shaped like q / k / v).  ``o`` (forward output) is optional:
    with it the kernel uses the one-pass D_i = dO_i . O_i form.  ``dq_dead_rows``: that many rows in front of dq's
    first row (the dead fusion-prefix query slots of a packed dqkv buffer) are zero-filled by the same launch."""
    B, Nq, H, dqk = q.shape
    Nk, dvd = k.shape[1], v.shape[3]
    a = _cabi.AttnBwdArgs()
    a.q, a.q_bs, a.q_rs = _bhs(q, "q")
    a.k, a.k_bs, a.k_rs = _bhs(k, "k")
    a.v, a.v_bs, a.v_rs = _bhs(v, "v")
    a.d_o, a.do_bs, a.do_rs = _bhs(d_o, "d_o")
    a.lse = _need(lse, torch.float32, "lse").data_ptr()
    a.dq, a.dq_bs, a.dq_rs = _bhs(dq, "dq")
    a.dk, a.dk_bs, a.dk_rs = _bhs(dk, "dk")
    a.dv_, a.dv_bs, a.dv_rs = _bhs(dv, "dv")
    a.B, a.H, a.Nq, a.Nk, a.dqk, a.dv = B, H, Nq, Nk, dqk, dvd
    a.scale, a.accumulate_dq = float(scale), int(accumulate_dq)
    a.dq_dead_rows = int(dq_dead_rows)
    if o is not None:
        a.o, a.o_bs, a.o_rs = _bhs(o, "o")
        assert a.o % 16 == 0 and a.o_bs % 8 == 0 and a.o_rs % 8 == 0, "o rows must be 16-byte aligned"
    check(_cabi.lib().davf_attention_bwd(C.byref(a), _stream()), "davf_attention_bwd")


# --------------------------------------------------------------------------------------------
# decoder assembly
# --------------------------------------------------------------------------------------------
def decoder_assemble_fwd(e: Tensor, ef: Tensor, mask_token: Tensor, pos: Tensor, ids_restore: Tensor, nK: int, nF: int) -> Tensor:
    """e f32 [B*nK,D], ef f32 [B*nF,D], mask_token f32 [D], pos f32 [L,D], ids_restore i64 [B,L] -> seq f32 [B,nF+L,D]."""
    B, L = ids_restore.shape
    D = e.shape[-1]
    for t, n in ((e, "e"), (ef, "ef"), (mask_token, "mask_token"), (pos, "pos")):
        _need(t, torch.float32, n)
    _need(ids_restore, torch.int64, "ids_restore")
    seq = torch.empty(B, nF + L, D, dtype=torch.float32, device=e.device)
    check(_cabi.lib().davf_decoder_assemble_fwd(_ptr(e), _ptr(ef), _ptr(mask_token), _ptr(pos), _ptr(ids_restore), _ptr(seq),
                                               B, nK, nF, L, D, _stream()), "davf_decoder_assemble_fwd")
    return seq


def decoder_assemble_bwd(dseq: Tensor, ids_keep: Tensor, ids_restore: Tensor, nF: int, dmask_token: Tensor, dpos: Tensor):
    """dseq f32 [B,nF+L,D] -> (de bf16 [B*nK,D], def bf16 [B*nF,D]); dmask_token [D] / dpos [L,D] accumulated."""
    _need(dseq, torch.float32, "dseq")
    B, S, D = dseq.shape
    L, nK = S - nF, ids_keep.shape[1]
    de = torch.empty(B * nK, D, dtype=torch.bfloat16, device=dseq.device)
    df = torch.empty(B * nF, D, dtype=torch.bfloat16, device=dseq.device)
    check(_cabi.lib().davf_decoder_assemble_bwd(_ptr(dseq), _ptr(_need(ids_keep, torch.int64, "ids_keep")),
                                               _ptr(_need(ids_restore, torch.int64, "ids_restore")), _ptr(de), _ptr(df),
                                               _ptr(_need(dmask_token, torch.float32, "dmask_token")), _ptr(_need(dpos, torch.float32, "dpos")),
                                               B, nK, nF, L, D, _stream()), "davf_decoder_assemble_bwd")
    return de, df


# --------------------------------------------------------------------------------------------
# loss
# --------------------------------------------------------------------------------------------
def masked_mse_fwd(img: Tensor, pred: Tensor, mask: Tensor, p: int, pred_G: int, pred_off: int, norm_pix: bool) -> Tensor:
    """Returns loss_sum f32 [1] = sum over masked patches of mean((pred - target)^2)."""
    _need(img, torch.float32, "img"); _need(pred, torch.float32, "pred"); _need(mask, torch.float32, "mask")
    B, Cc, H, W = img.shape
    out = torch.zeros(1, dtype=torch.float32, device=img.device)
    check(_cabi.lib().davf_masked_mse_fwd(_ptr(img), _ptr(pred), _ptr(mask), _ptr(out), B, Cc, H, W, p, pred_G, pred_off,
                                         int(norm_pix), _stream()), "davf_masked_mse_fwd")
    return out


def masked_mse_bwd(img: Tensor, pred: Tensor, mask: Tensor, gscale: Tensor, inv_count: float, p: int, pred_G: int,
                   pred_off: int, norm_pix: bool) -> Tensor:
    """Returns dpred bf16 [B*L, P]."""
    B, Cc, H, W = img.shape
    L = (H // p) * (W // p)
    dpred = torch.empty(B * L, p * p * Cc, dtype=torch.bfloat16, device=img.device)
    check(_cabi.lib().davf_masked_mse_bwd(_ptr(img), _ptr(pred), _ptr(mask), _ptr(_need(gscale, torch.float32, "gscale")),
                                         float(inv_count), _ptr(dpred), B, Cc, H, W, p, pred_G, pred_off, int(norm_pix), _stream()),
          "davf_masked_mse_bwd")
    return dpred


# --------------------------------------------------------------------------------------------
# optimizer
# --------------------------------------------------------------------------------------------
def adamw_step(p: Tensor, g: Tensor, m: Tensor, v: Tensor, p_bf16: Optional[Tensor], chunk_group: Tensor,
               hp: Tensor, scal: Tensor, beta1: float, beta2: float, eps: float, zero_grad: bool,
               sumsq_out: Optional[Tensor] = None) -> None:
    """Fused AdamW over the flat buffers; chunk_group u8 [n/64] (255 = frozen), hp f32 [ngroups*2],
    scal f32 [4] = {beta1^t, beta2^t, grad_scale, -}; all tables on the device."""
    for t, n in ((p, "p"), (g, "g"), (m, "m"), (v, "v"), (hp, "hp"), (scal, "scal")):
        _need(t, torch.float32, n)
    _need(chunk_group, torch.uint8, "chunk_group")
    assert chunk_group.numel() * 64 == p.numel()
    check(_cabi.lib().davf_adamw_step(_ptr(p), _ptr(g), _ptr(m), _ptr(v), _ptr(p_bf16), p.numel(), _ptr(chunk_group),
                                     _ptr(hp), _ptr(scal), beta1, beta2, eps, int(zero_grad), _ptr(sumsq_out), _stream()),
          "davf_adamw_step")


def sumsq_f32(g: Tensor, out: Tensor) -> None:
    check(_cabi.lib().davf_sumsq_f32(_ptr(_need(g, torch.float32, "g")), g.numel(), _ptr(_need(out, torch.float32, "out")), _stream()),
          "davf_sumsq_f32")


# --------------------------------------------------------------------------------------------
# a11 classifier tail (f32)
# --------------------------------------------------------------------------------------------
def meanpool_fwd(x: Tensor) -> Tensor:
    """x f32 [B, n, D] (dense rows, any batch stride) -> [B, D] mean over the tokens (classifier.py:49)."""
    _need(x, torch.float32, "x", contiguous=False)
    B, n, D = x.shape
    assert x.stride(2) == 1 and x.stride(1) == D
    out = torch.empty(B, D, dtype=torch.float32, device=x.device)
    check(_cabi.lib().davf_meanpool_fwd(_ptr(x), x.stride(0), B, n, D, _ptr(out), _stream()), "davf_meanpool_fwd")
    return out


def meanpool_bwd(dy: Tensor, n: int) -> Tensor:
    _need(dy, torch.float32, "dy")
    B, D = dy.shape
    dx = torch.empty(B, n, D, dtype=torch.float32, device=dy.device)
    check(_cabi.lib().davf_meanpool_bwd(_ptr(dy), B, n, D, _ptr(dx), _stream()), "davf_meanpool_bwd")
    return dx


def batchnorm1d_fwd(x: Tensor, running_mean: Tensor, running_var: Tensor, training: bool, momentum: float, eps: float):
    """BatchNorm1d(affine=False) on [B, D]; returns (y, save_mean, save_rstd); running stats updated in place when training."""
    _need(x, torch.float32, "x")
    B, D = x.shape
    y = torch.empty_like(x)
    mean = torch.empty(D, dtype=torch.float32, device=x.device)
    rstd = torch.empty(D, dtype=torch.float32, device=x.device)
    check(_cabi.lib().davf_batchnorm1d_fwd(_ptr(x), B, D, int(training), _ptr(_need(running_mean, torch.float32, "running_mean")),
                                           _ptr(_need(running_var, torch.float32, "running_var")), float(momentum), float(eps),
                                           _ptr(y), _ptr(mean), _ptr(rstd), _stream()), "davf_batchnorm1d_fwd")
    return y, mean, rstd


def batchnorm1d_bwd(dy: Tensor, x: Tensor, mean: Tensor, rstd: Tensor, training: bool) -> Tensor:
    _need(dy, torch.float32, "dy"); _need(x, torch.float32, "x")
    B, D = x.shape
    dx = torch.empty_like(x)
    check(_cabi.lib().davf_batchnorm1d_bwd(_ptr(dy), _ptr(x), _ptr(mean), _ptr(rstd), B, D, int(training), _ptr(dx), _stream()), "davf_batchnorm1d_bwd")
    return dx


def head_fwd(x: Tensor, W: Tensor, bias: Optional[Tensor]) -> Tensor:
    """y = x W^T + bias, f32, any number of classes."""
    _need(x, torch.float32, "x"); _need(W, torch.float32, "W")
    B, D = x.shape
    Cn = W.shape[0]
    y = torch.empty(B, Cn, dtype=torch.float32, device=x.device)
    check(_cabi.lib().davf_head_fwd(_ptr(x), _ptr(W), _ptr(bias), B, Cn, D, _ptr(y), _stream()), "davf_head_fwd")
    return y


def head_bwd(dy: Tensor, x: Tensor, W: Tensor, dW: Optional[Tensor], db: Optional[Tensor], need_dx: bool) -> Optional[Tensor]:
    """dW += dy^T x, db += colsum dy (in place, either may be None); returns dx = dy W or None."""
    _need(dy, torch.float32, "dy"); _need(x, torch.float32, "x"); _need(W, torch.float32, "W")
    B, D = x.shape
    Cn = W.shape[0]
    dx = torch.empty_like(x) if need_dx else None
    check(_cabi.lib().davf_head_bwd(_ptr(dy), _ptr(x), _ptr(W), B, Cn, D, _ptr(dW), _ptr(db), _ptr(dx), _stream()), "davf_head_bwd")
    return dx


# --------------------------------------------------------------------------------------------
# input stage: log-mel + frame normalisation
# --------------------------------------------------------------------------------------------
def logmel_workspace(sample_rate: int, n_fft: int, hop: int, n_mels: int, device) -> Tensor:
    """Device tables of the log-mel kernel (twiddles, window, mel filter bank), filled once."""
    ws = torch.empty(int(_cabi.lib().davf_logmel_workspace_bytes()), dtype=torch.uint8, device=device)
    check(_cabi.lib().davf_logmel_init(_ptr(ws), sample_rate, n_fft, hop, n_mels, _stream()), "davf_logmel_init")
    return ws


def logmel_fwd(ws: Tensor, wave: Tensor, gain_db: Optional[Tensor], n_mels: int, frames: int, eps: float) -> Tensor:
    """wave f32 or int16 PCM [B, T] -> log-mel f32 [B, 1, n_mels, frames]."""
    if not wave.is_cuda:
        raise RuntimeError("wave: expected a CUDA tensor (the sm_100a kernels have no CPU path)")
    assert wave.dim() == 2 and wave.is_contiguous() and wave.dtype in (torch.float32, torch.int16)
    B, T = wave.shape
    out = torch.empty(B, 1, n_mels, frames, dtype=torch.float32, device=wave.device)
    is16 = wave.dtype == torch.int16
    check(_cabi.lib().davf_logmel_fwd(_ptr(ws), _ptr(None if is16 else wave), _ptr(wave if is16 else None),
                                     _ptr(None if gain_db is None else _need(gain_db, torch.float32, "gain_db")), B, T, n_mels, frames, float(eps),
                                     _ptr(out), _stream()), "davf_logmel_fwd")
    return out


def image_normalize_u8(img: Tensor, mean: Sequence[float], std: Sequence[float]) -> Tensor:
    """uint8 [B, H, W, C] -> f32 [B, C, H, W] normalised."""
    _need(img, torch.uint8, "img")
    B, H, W, Cc = img.shape
    out = torch.empty(B, Cc, H, W, dtype=torch.float32, device=img.device)
    m = (C.c_float * Cc)(*[float(x) for x in mean])
    sd = (C.c_float * Cc)(*[float(x) for x in std])
    check(_cabi.lib().davf_image_normalize_u8(_ptr(img), _ptr(out), B, H, W, Cc, m, sd, _stream()), "davf_image_normalize_u8")
    return out


# --------------------------------------------------------------------------------------------
# stochastic depth (fine-tuning only)
# --------------------------------------------------------------------------------------------
def scale_rows_add(res: Tensor, y: Tensor, scale: Tensor, rows_per_sample: int) -> Tensor:
    """res + scale[sample] * y, f32 [rows, D] (rows sample-major)."""
    _need(res, torch.float32, "res"); _need(y, torch.float32, "y"); _need(scale, torch.float32, "scale")
    D = res.shape[-1]
    rows = res.numel() // D
    assert y.numel() == res.numel() and scale.numel() * rows_per_sample == rows
    out = torch.empty_like(res)
    check(_cabi.lib().davf_scale_rows_add(_ptr(res), _ptr(y), _ptr(scale), rows_per_sample, rows, D, _ptr(out), _stream()), "davf_scale_rows_add")
    return out


def scale_rows(src: Tensor, scale: Tensor, rows_per_sample: int, want_f32: bool = False, want_bf16: bool = True):
    """scale[sample] * src as (f32 or None, bf16 or None), both [rows, D]."""
    _need(src, torch.float32, "src"); _need(scale, torch.float32, "scale")
    D = src.shape[-1]
    rows = src.numel() // D
    assert scale.numel() * rows_per_sample == rows
    f = torch.empty(rows, D, dtype=torch.float32, device=src.device) if want_f32 else None
    b = torch.empty(rows, D, dtype=torch.bfloat16, device=src.device) if want_bf16 else None
    check(_cabi.lib().davf_scale_rows(_ptr(src), _ptr(scale), rows_per_sample, rows, D, _ptr(f), _ptr(b), _stream()), "davf_scale_rows")
    return f, b


def set_gemm_sms(n: int) -> int:
    """SMs the persistent GEMM grids use (see davf_set_gemm_sms); returns the value in effect."""
    return int(_cabi.lib().davf_set_gemm_sms(int(n)))


def launch_count() -> int:
    return int(_cabi.lib().davf_launch_count())


KIND_GEMM_2CTA, KIND_ATTN_TC, KIND_ATTN_MMA = 1, 2, 3


def launch_count_kind(kind: int) -> int:
    """Launches of one kernel family so far (see davf_launch_count_kind)."""
    return int(_cabi.lib().davf_launch_count_kind(int(kind)))


def set_pdl(on: bool) -> None:
    """Programmatic dependent launch on / off (see davf_set_pdl)."""
    check(_cabi.lib().davf_set_pdl(int(on)), "davf_set_pdl")


def set_gemm_2cta(on: bool) -> None:
    check(_cabi.lib().davf_set_gemm_2cta(int(on)), "davf_set_gemm_2cta")


def set_attn_impl(impl: int) -> None:
    check(_cabi.lib().davf_set_attn_impl(impl), "davf_set_attn_impl")

"""Hand-ordered forward / backward of the hot-path building blocks.

Each ``torch.autograd.Function`` below is one fused region of the model whose forward AND
backward are explicit sequences of libdavf_sm100 kernel launches (``kernels.*``); autograd only
stitches the regions together.  Parameter gradients never travel through autograd: wgrad GEMMs
and bias / LayerNorm reductions accumulate straight into the flat gradient buffer
(``ParamStore.grad``), so the Functions return ``None`` for every parameter.

Precision contract (bf16 path of BASELINE.json; the reference runs the same graph under autocast):
GEMM / attention operands bf16, accumulation f32; LayerNorm statistics, softmax, the residual
stream, the loss and all parameter gradients f32.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Optional, Tuple

import torch
from torch import nn

from . import kernels as K
from .params import ParamStore

Tensor = torch.Tensor


# --------------------------------------------------------------------------------------------
# linear helpers (forward: K-major x K-major; dgrad: K x MN; wgrad: MN x MN -- no transposes)
# --------------------------------------------------------------------------------------------
def _w2d(t: Tensor, cols: Optional[Tuple[int, int]]) -> Tensor:
    t = t.view(t.shape[0], -1)
    return t if cols is None else t[:, cols[0]:cols[1]]


def linear_fwd(st: ParamStore, x: Tensor, W: nn.Parameter, b: Optional[nn.Parameter], cols=None, **epi):
    """y = x W[:, cols]^T (+ b) with the fused epilogue options of kernels.gemm."""
    return K.gemm(x, _w2d(st.lowp(W), cols), True, True, bias=None if b is None else b.data, **epi)


def launch_wgrads(st: ParamStore, probs) -> None:
    """probs = [((dy, x, False, False), kwargs)]: weight-gradient GEMMs (+ bias-gradient row sums) of one backward
    region as ONE grouped launch.  Nothing in backward consumes dW / db, so the launch leaves the dgrad critical
    path: it is issued on a dedicated stream that is joined only at the end of backward."""
    if not probs:
        return
    ws = st.wgrad_stream()
    if ws is None:
        K.gemm_grouped(probs)
        return
    ws.wait_stream(torch.cuda.current_stream())
    for (dy, x, _, _), _kw in probs:
        dy.record_stream(ws)
        x.record_stream(ws)
    with torch.cuda.stream(ws):
        K.gemm_grouped(probs)


def linear_bwd(st: ParamStore, dy: Tensor, x: Tensor, W: nn.Parameter, b: Optional[nn.Parameter], cols=None,
               need_dx: bool = True, defer: Optional[list] = None, **epi):
    """Accumulates dW (+= dy^T x) and db (+= colsum dy) into the flat gradient buffer and returns
    dx = dy W[:, cols] (with optional fused epilogue) or None.  With ``defer`` (a list) the wgrad problem is appended
    to it instead of being launched: the caller groups a region's wgrads into one launch (``launch_wgrads``)."""
    want_b = b is not None and b.requires_grad
    if W.requires_grad:       # bias gradient rides along in the wgrad launch (ones-tile MMA)
        prob = ((dy, x, False, False), dict(out=_w2d(st.grad(W), cols), accumulate=True, rowsum_out=st.grad(b) if want_b else None))
        if defer is not None:
            defer.append(prob)
        else:
            launch_wgrads(st, [prob])
    elif want_b:
        K.colsum_bf16(dy, st.grad(b))
    if not need_dx:
        return None
    return K.gemm(dy, _w2d(st.lowp(W), cols), True, False, **epi)


class _Branches:
    """Independent launches of one region on auxiliary streams.  ``run(i, fn)`` forks stream i off the region's stream,
    calls ``fn`` there and marks the tensors it returns as used by the region's stream; ``join()`` makes the region's
    stream wait for every branch that ran.  Without auxiliary streams (CPU tensors, DAVF_STREAMS=0) everything runs inline."""

    def __init__(self, st: ParamStore):
        self.aux = st.aux_streams(2)
        self.cur = torch.cuda.current_stream() if self.aux else None
        self.used = set()

    def run(self, i: int, fn):
        if not self.aux:
            return fn()
        s = self.aux[i % len(self.aux)]
        if i % len(self.aux) not in self.used:
            s.wait_stream(self.cur)
            self.used.add(i % len(self.aux))
        with torch.cuda.stream(s):
            out = fn()
        for t in (out if isinstance(out, (tuple, list)) else (out,)):
            if isinstance(t, torch.Tensor):
                t.record_stream(self.cur)
        return out

    def join(self):
        if self.aux:
            for i in self.used:
                self.cur.wait_stream(self.aux[i])
            self.used = set()


def _f32_rows(t: Tensor) -> Tensor:
    return t.reshape(-1, t.shape[-1])


def _lowp_grad(st: ParamStore, dy2d: Tensor) -> Tensor:
    """bf16 copy of an incoming f32 gradient [rows, D]: the one its producer already wrote, else a cast launch."""
    lp = st.take_lowp_grad(dy2d)
    if lp is None:
        return K.cast_rows_bf16(dy2d)
    if lp.is_cuda:       # written on the producer's stream, read by this region's GEMMs on the current one
        lp.record_stream(torch.cuda.current_stream())
    return lp.view(dy2d.shape)


def _new_lowp(like: Tensor) -> Tensor:
    return torch.empty(like.numel() // like.shape[-1], like.shape[-1], dtype=torch.bfloat16, device=like.device)


def droppath_scale(B: int, drop_prob: float, device) -> Tensor:
    """Per-sample keep-mask / keep_prob of timm DropPath (scale_by_keep=True).  The RNG draw cannot reproduce the
    reference's ``bernoulli_`` stream bit for bit (parity runs use drop_path = 0 or inject the masks)."""
    keep = 1.0 - drop_prob
    return (torch.rand(B, device=device) < keep).to(torch.float32) / keep


def params_of(ns: SimpleNamespace):
    """All nn.Parameters referenced by a region namespace (nested namespaces included)."""
    out = []
    for v in vars(ns).values():
        if isinstance(v, nn.Parameter):
            out.append(v)
        elif isinstance(v, SimpleNamespace):
            out.extend(params_of(v))
    return out


def _done(m: SimpleNamespace) -> None:
    pl = m.__dict__.get("_plist")
    if pl is None:
        pl = m.__dict__["_plist"] = params_of(m)
    m.store.done(pl)


# --------------------------------------------------------------------------------------------
# a2  patch embed (+pos, kept rows only)        vits.py:91-100, timm PatchEmbed
# --------------------------------------------------------------------------------------------
class PatchEmbedFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, image: Tensor, ids_keep: Optional[Tensor], weight: nn.Parameter, m: SimpleNamespace):
        st: ParamStore = m.store
        B = image.shape[0]
        p = m.patch
        L = (image.shape[2] // p) * (image.shape[3] // p)
        a = K.patch_rows(image, ids_keep, p)                       # [B*nK, C*p*p] bf16, kept patches only
        nK = a.shape[0] // B
        if ids_keep is None:
            idx = torch.arange(L, device=image.device, dtype=torch.int64).repeat(B)
        else:
            idx = ids_keep.reshape(-1)
        D = weight.shape[0]
        x = linear_fwd(st, a, weight, m.bias, res=m.pos_embed.data.view(L, D), res_idx=idx, out_dtype=torch.float32)
        ctx.m = m
        ctx.save_for_backward(a)
        return x.view(B, nK, D)

    @staticmethod
    def backward(ctx, dx: Tensor):
        m = ctx.m
        (a,) = ctx.saved_tensors
        dxb = _lowp_grad(m.store, _f32_rows(dx.contiguous()))
        linear_bwd(m.store, dxb, a, m.weight, m.bias, need_dx=False)
        _done(m)
        return None, None, None, None


# --------------------------------------------------------------------------------------------
# a5  fusion tokens broadcast over the batch (deepavfusion.py:97): backward = batch sum
# --------------------------------------------------------------------------------------------
class BroadcastTokensFn(torch.autograd.Function):
    """``fusion_tokens.expand(B, -1, -1)`` with the gradient reduced by our kernel straight into the flat
    gradient buffer, so that NO parameter gradient travels through autograd ``AccumulateGrad`` nodes
    (those pin the stream they were first created on, which breaks multi-stream CUDA-graph capture)."""

    @staticmethod
    def forward(ctx, tokens: Tensor, B: int, m: SimpleNamespace):
        ctx.m = m
        return tokens.detach().expand(B, -1, -1).contiguous()

    @staticmethod
    def backward(ctx, dx: Tensor):
        m = ctx.m
        st: ParamStore = m.store
        dx = dx.contiguous()
        _, F, D = dx.shape
        if m.tokens.requires_grad:
            K.batchsum_f32(dx, 0, F, st.grad(m.tokens).view(F, D), True)
        _done(m)
        return None, None, None


# --------------------------------------------------------------------------------------------
# a5  a tensor with several consumers (deepavfusion.py:104-106): backward = ONE fan-in kernel
# --------------------------------------------------------------------------------------------
class FanOutFn(torch.autograd.Function):
    """``x`` feeds ``n`` consumers (x_image / x_audio: their own block and the fusion block; x_fusion: all three blocks).
    Forward returns ``n`` aliases; backward sums the consumers' gradients in one kernel that also writes the bf16 copy
    the upstream region needs as a GEMM operand -- instead of autograd's n - 1 add kernels plus a cast."""

    @staticmethod
    def forward(ctx, x: Tensor, n: int, st: ParamStore):
        ctx.st = st
        return tuple(x.view_as(x) for _ in range(n))

    @staticmethod
    def backward(ctx, *grads):
        parts = [g.contiguous() for g in grads if g is not None]
        if not parts:
            return None, None, None
        if len(parts) == 1:
            return parts[0], None, None
        out, lp = K.sum_cast(parts)
        ctx.st.stash_lowp_grad(out, lp)
        return out, None, None


# --------------------------------------------------------------------------------------------
# a3  attention half of a timm Block:  y = x + proj(attn(qkv(LN(cat(xp, x)))))  on the live rows
# --------------------------------------------------------------------------------------------
class AttnBranchFn(torch.autograd.Function):
    """xp = optional prefix rows (the fusion tokens of deepavfusion.py:104-105) that act as
    keys / values only: their block outputs are discarded by the reference, so no query / proj /
    MLP work is done for them."""

    @staticmethod
    def forward(ctx, xp: Optional[Tensor], x: Tensor, anchor: Tensor, m: SimpleNamespace, drop: Optional[Tensor] = None):
        """``drop``: per-sample DropPath scale [B] of this residual branch (None: no stochastic depth)."""
        st: ParamStore = m.store
        B, n, D = x.shape
        x = x.contiguous()
        if xp is not None:
            xp = xp.contiguous()
            nP = xp.shape[1]
            x0, x1 = xp, x
        else:
            nP = 0
            x0, x1 = x, None
        S = nP + n
        H = m.heads
        hd = D // H
        xn, _, mean, rstd = K.layernorm_fwd(x0, x1, m.norm_w.data, m.norm_b.data, m.eps)
        qkv = linear_fwd(st, xn, m.qkv_w, m.qkv_b)                                    # [B*S, 3D] bf16
        q5 = qkv.view(B, S, 3, H, hd)
        o, lse = K.attention_fwd(q5[:, nP:, 0], q5[:, :, 1], q5[:, :, 2], hd ** -0.5)  # [B,n,H,hd]
        if drop is None:
            y = linear_fwd(st, o.view(B * n, D), m.proj_w, m.proj_b, res=x.view(B * n, D), out_dtype=torch.float32)
        else:
            y = K.scale_rows_add(x.view(B * n, D), linear_fwd(st, o.view(B * n, D), m.proj_w, m.proj_b, out_dtype=torch.float32), drop, n)
        ctx.m, ctx.nP, ctx.drop = m, nP, drop
        ctx.save_for_backward(x0, x1, mean, rstd, xn, qkv, o, lse)
        return y.view(B, n, D)

    @staticmethod
    def backward(ctx, dy: Tensor):
        m, nP = ctx.m, ctx.nP
        st: ParamStore = m.store
        x0, x1, mean, rstd, xn, qkv, o, lse = ctx.saved_tensors
        dy = dy.contiguous()
        B, n, D = dy.shape
        S, H = nP + n, m.heads
        hd = D // H
        dyb = _lowp_grad(st, dy.view(B * n, D)) if ctx.drop is None else K.scale_rows(dy.view(B * n, D), ctx.drop, n)[1]
        wg = []                                                                       # proj + qkv wgrads: one grouped launch
        do = linear_bwd(st, dyb, o.view(B * n, D), m.proj_w, m.proj_b, defer=wg)      # [B*n, D] bf16
        dqkv = torch.empty_like(qkv)
        d5 = dqkv.view(B, S, 3, H, hd)
        q5 = qkv.view(B, S, 3, H, hd)
        K.attention_bwd(q5[:, nP:, 0], q5[:, :, 1], q5[:, :, 2], do.view(B, n, H, hd), lse, hd ** -0.5,
                        d5[:, nP:, 0], d5[:, :, 1], d5[:, :, 2], o=o, dq_dead_rows=nP)      # (also zero-fills the dead query slots)
        dxn = linear_bwd(st, dqkv, xn, m.qkv_w, m.qkv_b, defer=wg)                    # [B*S, D] bf16
        launch_wgrads(st, wg)
        gw, gb = st.grad(m.norm_w), st.grad(m.norm_b)
        if nP:
            dxp, dx = K.layernorm_bwd(x0, x1, m.norm_w.data, mean, rstd, dxn, None, None, dy, gw, gb,
                                      need_dx0=ctx.needs_input_grad[0])
        else:       # single consumer upstream (the previous block's MLP branch): hand it the bf16 copy too
            lp = _new_lowp(dy)
            dx, _ = K.layernorm_bwd(x0, None, m.norm_w.data, mean, rstd, dxn, None, dy, None, gw, gb, dx0_lowp=lp)
            st.stash_lowp_grad(dx, lp)
            dxp = None
        _done(m)
        return dxp, dx, None, None, None


# --------------------------------------------------------------------------------------------
# MLP half of a timm Block / fusion block:  y = x + fc2(gelu(fc1(LN(x))))
# --------------------------------------------------------------------------------------------
class MlpBranchFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x: Tensor, anchor: Tensor, m: SimpleNamespace, drop: Optional[Tensor] = None):
        st: ParamStore = m.store
        x = x.contiguous()
        D = x.shape[-1]
        x2 = x.view(1, -1, D)
        xn, _, mean, rstd = K.layernorm_fwd(x2, None, m.norm_w.data, m.norm_b.data, m.eps)
        a, h = linear_fwd(st, xn, m.fc1_w, m.fc1_b, act=K.ACT_GELU, want_aux=True)   # a = gelu(z), h = gelu'(z) (bf16), z = fc1 output
        if drop is None:
            y = linear_fwd(st, a, m.fc2_w, m.fc2_b, res=x2.view(-1, D), out_dtype=torch.float32)
        else:                                                                         # DropPath: x + scale[b] * branch
            y = K.scale_rows_add(x2.view(-1, D), linear_fwd(st, a, m.fc2_w, m.fc2_b, out_dtype=torch.float32), drop, x.shape[1])
        ctx.m, ctx.drop, ctx.rps = m, drop, x.shape[1]
        ctx.save_for_backward(x2, mean, rstd, xn, h, a)
        return y.view(x.shape)

    @staticmethod
    def backward(ctx, dy: Tensor):
        m = ctx.m
        st: ParamStore = m.store
        x2, mean, rstd, xn, h, a = ctx.saved_tensors
        dy = dy.contiguous()
        D = dy.shape[-1]
        dyb = _lowp_grad(st, dy.view(-1, D)) if ctx.drop is None else K.scale_rows(dy.view(-1, D), ctx.drop, ctx.rps)[1]
        wg = []                                                                       # fc2 + fc1 wgrads: one grouped launch
        dh = linear_bwd(st, dyb, a, m.fc2_w, m.fc2_b, defer=wg, act=K.ACT_DGELU, aux_in=h)   # dgrad times the saved gelu'
        dxn = linear_bwd(st, dh, xn, m.fc1_w, m.fc1_b, defer=wg)
        launch_wgrads(st, wg)
        lp = _new_lowp(dy)                                                            # consumed by the attention branch's backward
        dx, _ = K.layernorm_bwd(x2, None, m.norm_w.data, mean, rstd, dxn, None, dy.view(1, -1, D), None,
                                st.grad(m.norm_w), st.grad(m.norm_b), dx0_lowp=lp)
        dx = dx.view(dy.shape)
        st.stash_lowp_grad(dx, lp)
        _done(m)
        return dx, None, None, None


# --------------------------------------------------------------------------------------------
# final norms (vits.py:116 / deepavfusion.py:111-113): f32 in, f32 out
# --------------------------------------------------------------------------------------------
class LayerNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x: Tensor, anchor: Tensor, m: SimpleNamespace):
        x = x.contiguous()
        D = x.shape[-1]
        x2 = x.view(1, -1, D)
        _, y, mean, rstd = K.layernorm_fwd(x2, None, m.norm_w.data, m.norm_b.data, m.eps, want_bf16=False, want_f32=True)
        ctx.m = m
        ctx.save_for_backward(x2, mean, rstd)
        return y.view(x.shape)

    @staticmethod
    def backward(ctx, dy: Tensor):
        m = ctx.m
        st: ParamStore = m.store
        x2, mean, rstd = ctx.saved_tensors
        dy = dy.contiguous()
        lp = _new_lowp(dy)
        dx, _ = K.layernorm_bwd(x2, None, m.norm_w.data, mean, rstd, None, dy.view(-1, dy.shape[-1]), None, None,
                                st.grad(m.norm_w), st.grad(m.norm_b), dx0_lowp=lp)
        dx = dx.view(dy.shape)
        st.stash_lowp_grad(dx, lp)
        _done(m)
        return dx, None, None


# --------------------------------------------------------------------------------------------
# a4  fusion-token attention of FusionBlock_FactorizedAVInteractions (fusion_blocks.py:235-283)
#     out = LN_mm(xmm) + cat(proj(pair_attn), attn_v, attn_a)
# --------------------------------------------------------------------------------------------
class _Lin:
    """One Linear of a grouped launch, described by raw views: ``w`` bf16 [N, K] (possibly a column slice of a
    stacked weight), ``b`` f32 bias or None, ``gw`` / ``gb`` the gradient views they accumulate into (None =
    frozen)."""
    __slots__ = ("w", "b", "gw", "gb")

    def __init__(self, w, b=None, gw=None, gb=None):
        self.w, self.b, self.gw, self.gb = w, b, gw, gb


def _lin_of(st: ParamStore, W: nn.Parameter, b: Optional[nn.Parameter]) -> "_Lin":
    return _Lin(st.lowp(W), None if b is None else b.data, st.grad(W) if W.requires_grad else None,
                st.grad(b) if (b is not None and b.requires_grad) else None)


def group_fwd(items):
    """items = [(x, _Lin, epilogue kwargs)]: independent forward Linears in one launch."""
    return K.gemm_grouped([((x, l.w, True, True), dict(bias=l.b, **kw)) for x, l, kw in items])


def group_bwd(st: ParamStore, items, defer: Optional[list] = None):
    """items = [(dy, x, _Lin, dgrad epilogue kwargs or None)]: the dgrad launches go out as one grouped launch on the
    current stream; the wgrad (+ bias gradient) problems are appended to ``defer`` (flushed by the caller with
    ``launch_wgrads`` before the region reports its parameters done) or, without it, launched as one group on the
    wgrad stream.  Returns the dx list (None where no dgrad was requested)."""
    wg = [((dy, x, False, False), dict(out=l.gw, accumulate=True, rowsum_out=l.gb)) for dy, x, l, _ in items if l.gw is not None]
    for dy, x, l, _ in items:
        if l.gw is None and l.gb is not None:
            K.colsum_bf16(dy, l.gb)
    if defer is not None:
        defer.extend(wg)
    else:
        launch_wgrads(st, wg)
    dg = [(i, ((dy, l.w, True, False), kw)) for i, (dy, x, l, kw) in enumerate(items) if kw is not None]
    out = [None] * len(items)
    for (i, _), r in zip(dg, K.gemm_grouped([c for _, c in dg])):
        out[i] = r
    return out


class FusionAttnFn(torch.autograd.Function):
    """The dense pair attention (:245-258) is evaluated in its exactly-equivalent factorised form
    (SURVEY.md 7.1-2): k(xva_ij) = Wk1 v_i + Wk2 a_j + bk  =>  softmax over the 8x8 pairs is the
    outer product of two 8-way softmaxes, and the value sum splits the same way; xva is never
    materialised and the k / v projections shrink 8x.

    Launch structure: the block's ~20 small Linears have M = 512 .. 3136 rows and are latency-bound one by
    one, and the fusion block is the critical path of every encoder layer (the modality blocks run beside it
    on other streams).  Independent Linears therefore share grouped launches (``K.gemm_grouped``), and the
    pair-attention k / v Linears -- same input, weights stacked in the flat buffer -- are one GEMM."""

    @staticmethod
    def _idx(m, B, F, off, n_tok, device):
        """Row indices of out[b, off:off+n_tok] in the [B*F, D] layout (cached: no per-step arange kernels)."""
        cache = m.__dict__.setdefault("_idx_cache", {})
        key = (B, F, off, n_tok, str(device))
        t = cache.get(key)
        if t is None:
            t = (torch.arange(B, device=device).view(B, 1) * F + off + torch.arange(n_tok, device=device).view(1, n_tok)).reshape(-1)
            cache[key] = t
        return t

    @staticmethod
    def _lins(m):
        st: ParamStore = m.store
        D = m.v_w.shape[0]
        kv_lp, kv_g = st.stacked(m.k_w, m.v_w)                       # [qk + D, 2D]
        train = m.k_w.requires_grad
        assert m.v_w.requires_grad == train
        if m.k_b is not None:
            kvb_lp, kvb_g = st.stacked(m.k_b, m.v_b)                 # bf16 shadow unused: the bias is read in f32
            k0, _ = st.span(st.index_of(m.k_b))
            kvb = st.flat_p[k0:k0 + kvb_g.numel()]
            kvb_g = kvb_g if m.k_b.requires_grad else None
        else:
            kvb, kvb_g = None, None
        pair_v = _Lin(kv_lp[:, :D], kvb, kv_g[:, :D] if train else None, kvb_g)      # bias on the v side only
        pair_a = _Lin(kv_lp[:, D:], None, kv_g[:, D:] if train else None, None)
        return SimpleNamespace(
            q_v=_lin_of(st, m.attn_v.q_w, m.attn_v.q_b), kv_v=_lin_of(st, m.attn_v.kv_w, m.attn_v.kv_b), proj_v=_lin_of(st, m.attn_v.proj_w, m.attn_v.proj_b),
            q_a=_lin_of(st, m.attn_a.q_w, m.attn_a.q_b), kv_a=_lin_of(st, m.attn_a.kv_w, m.attn_a.kv_b), proj_a=_lin_of(st, m.attn_a.proj_w, m.attn_a.proj_b),
            q2=_lin_of(st, m.q_w, m.q_b), pair_v=pair_v, pair_a=pair_a, proj=_lin_of(st, m.proj_w, m.proj_b))

    @staticmethod
    def forward(ctx, xmm: Tensor, xv: Tensor, xa: Tensor, anchor: Tensor, m: SimpleNamespace, drop: Optional[Tensor] = None):
        xmm, xv, xa = xmm.contiguous(), xv.contiguous(), xa.contiguous()
        B, F, D = xmm.shape
        nmm, nv, na = m.tkns
        H = m.heads
        hd = D // H
        qk = m.q_w.shape[0]
        dq = qk // H
        scale = hd ** -0.5                                                          # fusion_blocks.py:220-222
        L = FusionAttnFn._lins(m)
        seg = [0, nmm, nmm + nv, F]
        br = _Branches(m.store)           # independent small launches of the block run as parallel branches
        xv_n, _, mean_v, rstd_v = br.run(0, lambda: K.layernorm_fwd(xv, None, m.n_img_w.data, m.n_img_b.data, m.eps))
        xa_n, _, mean_a, rstd_a = br.run(1, lambda: K.layernorm_fwd(xa, None, m.n_aud_w.data, m.n_aud_b.data, m.eps))
        mm_b, mm_f, mean_m, rstd_m = K.layernorm_fwd(xmm, None, m.n_mm_w.data, m.n_mm_b.data, m.eps, True, True, seg)
        m2, mv, ma = mm_b[:B * nmm], mm_b[B * nmm:B * (nmm + nv)], mm_b[B * (nmm + nv):]
        br.join()
        out = torch.empty(B * F, D, dtype=torch.float32, device=xmm.device)
        res = mm_f if drop is None else None        # DropPath: the three projections land in `out` alone, the residual is added scaled
        Nv, Na = xv_n.shape[0] // B, xa_n.shape[0] // B

        # every Linear that only needs the normed inputs: CrossAttention q / kv of both modalities (:46-52) + pair q (:252)
        qv, kvv, qa, kva, q2 = group_fwd([(mv, L.q_v, {}), (xv_n, L.kv_v, {}), (ma, L.q_a, {}), (xa_n, L.kv_a, {}), (m2, L.q2, {})])
        kvv5, kva5 = kvv.view(B, Nv, 2, H, hd), kva.view(B, Na, 2, H, hd)
        oa, lse_a = br.run(0, lambda: K.attention_fwd(qa.view(B, na, H, hd), kva5[:, :, 0], kva5[:, :, 1], scale))
        ov, lse_v = K.attention_fwd(qv.view(B, nv, H, hd), kvv5[:, :, 0], kvv5[:, :, 1], scale)
        br.join()
        # out[b, off:off+n] = LN_mm(xmm)[b, off:...] + proj(.) ; the bf16 proj outputs feed the pair attention
        (_, pv), (_, pa) = group_fwd([
            (ov.view(B * nv, D), L.proj_v, dict(want_aux=True, res=res, out=out, window=(nv, F, nmm))),
            (oa.view(B * na, D), L.proj_a, dict(want_aux=True, res=res, out=out, window=(na, F, nmm + nv)))])
        # factorised pair attention: [k | v] of each side in one GEMM against the stacked weight
        kv2v, kv2a = group_fwd([(pv, L.pair_v, {}), (pa, L.pair_a, {})])            # [B*nv, qk + D], [B*na, qk + D]
        q2v = q2.view(B, nmm, H, dq)
        k_v, v_v = kv2v.view(B, nv, qk + D)[:, :, :qk].unflatten(2, (H, dq)), kv2v.view(B, nv, qk + D)[:, :, qk:].unflatten(2, (H, hd))
        k_a, v_a = kv2a.view(B, na, qk + D)[:, :, :qk].unflatten(2, (H, dq)), kv2a.view(B, na, qk + D)[:, :, qk:].unflatten(2, (H, hd))
        # the two halves add into one zero-initialised buffer with bf16x2 atomics, so they need no order
        o2 = torch.zeros(B, nmm, H, hd, dtype=torch.bfloat16, device=xmm.device)
        _, lse2a = br.run(0, lambda: K.attention_fwd(q2v, k_a, v_a, scale, out=o2, accumulate=2))
        _, lse2v = K.attention_fwd(q2v, k_v, v_v, scale, out=o2, accumulate=2)
        br.join()
        o2 = o2.view(B * nmm, D)
        K.gemm(o2, L.proj.w, bias=L.proj.b, res=res, out=out, window=(nmm, F, 0))
        if drop is not None:                         # xmm + drop_path(res_fusion), fusion_blocks.py:283
            out = K.scale_rows_add(mm_f, out, drop, F)
        ctx.m, ctx.drop = m, drop
        ctx.save_for_backward(xmm, xv, xa, mean_m, rstd_m, mean_v, rstd_v, mean_a, rstd_a, mm_b, xv_n, xa_n,
                              qv, kvv, ov, lse_v, pv, qa, kva, oa, lse_a, pa,
                              q2, kv2v, kv2a, lse2v, lse2a, o2)
        return out.view(B, F, D)

    @staticmethod
    def backward(ctx, dout: Tensor):
        m = ctx.m
        st: ParamStore = m.store
        (xmm, xv, xa, mean_m, rstd_m, mean_v, rstd_v, mean_a, rstd_a, mm_b, xv_n, xa_n,
         qv, kvv, ov, lse_v, pv, qa, kva, oa, lse_a, pa,
         q2, kv2v, kv2a, lse2v, lse2a, o2) = ctx.saved_tensors
        dout = dout.contiguous()
        B, F, D = dout.shape
        nmm, nv, na = m.tkns
        H = m.heads
        hd = D // H
        qk = m.q_w.shape[0]
        dq = qk // H
        scale = hd ** -0.5
        L = FusionAttnFn._lins(m)
        Nv, Na = xv_n.shape[0] // B, xa_n.shape[0] // B
        d_res = dout.view(B * F, D)                                                 # gradient of the residual path (unscaled)
        d2 = d_res if ctx.drop is None else K.scale_rows(d_res, ctx.drop, F, want_f32=True, want_bf16=False)[0]   # of the branch
        m2, mv, ma = mm_b[:B * nmm], mm_b[B * nmm:B * (nmm + nv)], mm_b[B * (nmm + nv):]
        dseg = torch.empty_like(mm_b)                                               # d LN_mm(xmm), segment-major bf16
        dm2, dmv, dma = dseg[:B * nmm], dseg[B * nmm:B * (nmm + nv)], dseg[B * (nmm + nv):]

        # ---- pair attention ----
        dr2 = K.cast_rows_bf16(d2, B * nmm, nmm, F, 0)
        wg = []                                      # the block's ten wgrads leave in two grouped launches at the end
        (do2,) = group_bwd(st, [(dr2, o2, L.proj, {})], wg)
        do2 = do2.view(B, nmm, H, hd)
        br = _Branches(st)
        dq2 = torch.zeros_like(q2)                   # both halves of the pair attention add into it (bf16x2 atomics)
        dkv2v, dkv2a = torch.empty_like(kv2v), torch.empty_like(kv2a)

        def split(t, n):
            t3 = t.view(B, n, qk + D)
            return t3[:, :, :qk].unflatten(2, (H, dq)), t3[:, :, qk:].unflatten(2, (H, hd))
        q2v = q2.view(B, nmm, H, dq)
        (k_v, v_v), (k_a, v_a) = split(kv2v, nv), split(kv2a, na)
        (dk_v, dv_v), (dk_a, dv_a) = split(dkv2v, nv), split(dkv2a, na)
        br.run(0, lambda: K.attention_bwd(q2v, k_a, v_a, do2, lse2a, scale, dq2.view(B, nmm, H, dq), dk_a, dv_a, accumulate_dq=2))
        K.attention_bwd(q2v, k_v, v_v, do2, lse2v, scale, dq2.view(B, nmm, H, dq), dk_v, dv_v, accumulate_dq=2)
        br.join()
        # d proj-output of each aggregation group = residual path (rows of dout) + [dk | dv] through the stacked weight
        idx_v = FusionAttnFn._idx(m, B, F, nmm, nv, dout.device)
        idx_a = FusionAttnFn._idx(m, B, F, nmm + nv, na, dout.device)
        _, dpv, dpa = group_bwd(st, [(dq2, m2, L.q2, dict(out=dm2)),
                                     (dkv2v, pv, L.pair_v, dict(res=d2, res_idx=idx_v)),
                                     (dkv2a, pa, L.pair_a, dict(res=d2, res_idx=idx_a))], wg)

        # ---- the two cross attentions ----
        do_v, do_a = group_bwd(st, [(dpv, ov.view(B * nv, D), L.proj_v, {}), (dpa, oa.view(B * na, D), L.proj_a, {})], wg)
        dqv, dkvv, dqa, dkva = torch.empty_like(qv), torch.empty_like(kvv), torch.empty_like(qa), torch.empty_like(kva)
        kvv5, dkvv5 = kvv.view(B, Nv, 2, H, hd), dkvv.view(B, Nv, 2, H, hd)
        kva5, dkva5 = kva.view(B, Na, 2, H, hd), dkva.view(B, Na, 2, H, hd)
        br.run(0, lambda: K.attention_bwd(qa.view(B, na, H, hd), kva5[:, :, 0], kva5[:, :, 1], do_a.view(B, na, H, hd), lse_a, scale,
                                          dqa.view(B, na, H, hd), dkva5[:, :, 0], dkva5[:, :, 1], o=oa.view(B, na, H, hd)))
        K.attention_bwd(qv.view(B, nv, H, hd), kvv5[:, :, 0], kvv5[:, :, 1], do_v.view(B, nv, H, hd), lse_v, scale,
                        dqv.view(B, nv, H, hd), dkvv5[:, :, 0], dkvv5[:, :, 1], o=ov.view(B, nv, H, hd))
        br.join()
        _, _, dxv_n, dxa_n = group_bwd(st, [(dqv, mv, L.q_v, dict(out=dmv)), (dqa, ma, L.q_a, dict(out=dma)),
                                            (dkvv, xv_n, L.kv_v, {}), (dkva, xa_n, L.kv_a, {})], wg)
        launch_wgrads(st, wg)

        gw_v, gb_v, gw_a, gb_a = st.grad(m.n_img_w), st.grad(m.n_img_b), st.grad(m.n_aud_w), st.grad(m.n_aud_b)
        dxv, _ = br.run(0, lambda: K.layernorm_bwd(xv, None, m.n_img_w.data, mean_v, rstd_v, dxv_n, None, None, None, gw_v, gb_v))
        dxa, _ = br.run(1, lambda: K.layernorm_bwd(xa, None, m.n_aud_w.data, mean_a, rstd_a, dxa_n, None, None, None, gw_a, gb_a))
        seg = [0, nmm, nmm + nv, F]
        dxmm, _ = K.layernorm_bwd(xmm, None, m.n_mm_w.data, mean_m, rstd_m, dseg, d_res, None, None,
                                  st.grad(m.n_mm_w), st.grad(m.n_mm_b), seg_start=seg)
        br.join()
        _done(m)
        return dxmm, dxv, dxa, None, None, None


# --------------------------------------------------------------------------------------------
# a6  decoder front: embed, mask tokens, unshuffle, +pos, cat fusion      avmae.py:158-169
# --------------------------------------------------------------------------------------------
class DecoderEmbedFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x: Tensor, xf: Tensor, ids_restore: Tensor, ids_keep: Tensor, anchor: Tensor, m: SimpleNamespace):
        st: ParamStore = m.store
        B, nK, D = x.shape
        nF = xf.shape[1]
        L = ids_restore.shape[1]
        Dd = m.embed_w.shape[0]
        xb = K.cast_rows_bf16(x.contiguous().view(B * nK, D))
        xfb = K.cast_rows_bf16(xf.contiguous().view(B * nF, D))
        e = linear_fwd(st, xb, m.embed_w, m.embed_b, out_dtype=torch.float32)
        ef = linear_fwd(st, xfb, m.embed_w, m.embed_b, out_dtype=torch.float32)    # same Linear (avmae.py:158)
        seq = K.decoder_assemble_fwd(e, ef, m.mask_token.data.view(Dd), m.pos_embed.data.view(L, Dd), ids_restore, nK, nF)
        ctx.m, ctx.nF = m, nF
        ctx.save_for_backward(xb, xfb, ids_restore, ids_keep)
        return seq

    @staticmethod
    def backward(ctx, dseq: Tensor):
        m, nF = ctx.m, ctx.nF
        st: ParamStore = m.store
        xb, xfb, ids_restore, ids_keep = ctx.saved_tensors
        B, L = ids_restore.shape
        nK = ids_keep.shape[1]
        Dd = m.embed_w.shape[0]
        D = xb.shape[1]
        de, df = K.decoder_assemble_bwd(dseq.contiguous(), ids_keep, ids_restore, nF,
                                        st.grad(m.mask_token).view(Dd), st.grad(m.pos_embed).view(L, Dd))
        dx = linear_bwd(st, de, xb, m.embed_w, m.embed_b, out_dtype=torch.float32)
        dxf = linear_bwd(st, df, xfb, m.embed_w, m.embed_b, out_dtype=torch.float32)
        _done(m)
        return dx.view(B, nK, D), dxf.view(B, nF, D), None, None, None, None


# --------------------------------------------------------------------------------------------
# a6/a7  decoder head + loss: pred = Linear(LN(seq[:, nF:]));  loss = masked normalised MSE
#        avmae.py:172,179 + :183-214
# --------------------------------------------------------------------------------------------
class PredLossFn(torch.autograd.Function):
    """Returns (loss, pred).  ``pred`` is returned for inspection / logging and is NOT
    differentiable (train.py:164 only consumes the two losses)."""

    @staticmethod
    def forward(ctx, seq: Tensor, img: Tensor, mask: Tensor, anchor: Tensor, m: SimpleNamespace):
        st: ParamStore = m.store
        seq = seq.contiguous()
        B, S, Dd = seq.shape
        L = mask.shape[1]
        nF = S - L
        xn, _, mean, rstd = K.layernorm_fwd(seq[:, nF:], None, m.norm_w.data, m.norm_b.data, m.eps)
        pred = linear_fwd(st, xn, m.pred_w, m.pred_b, out_dtype=torch.float32)      # [B*L, P]
        loss_sum = K.masked_mse_fwd(img, pred, mask, m.patch, L, 0, m.norm_pix)
        count = B * (L - m.len_keep)                                                # == mask.sum(), avmae.py:197
        ctx.m, ctx.count, ctx.nF = m, count, nF
        ctx.save_for_backward(seq, img, mask, mean, rstd, xn, pred)
        pred3 = pred.view(B, L, -1)
        ctx.mark_non_differentiable(pred3)
        return (loss_sum / count).reshape(()), pred3

    @staticmethod
    def backward(ctx, dloss: Tensor, _dpred):
        m, count, nF = ctx.m, ctx.count, ctx.nF
        st: ParamStore = m.store
        seq, img, mask, mean, rstd, xn, pred = ctx.saved_tensors
        B, S, Dd = seq.shape
        L = S - nF
        g = dloss.reshape(1).to(torch.float32).contiguous()
        dpred = K.masked_mse_bwd(img, pred, mask, g, 1.0 / count, m.patch, L, 0, m.norm_pix)
        dxn = linear_bwd(st, dpred, xn, m.pred_w, m.pred_b)
        dseq = torch.empty_like(seq)
        dseq[:, :nF].zero_()
        K.layernorm_bwd(seq[:, nF:], None, m.norm_w.data, mean, rstd, dxn, None, None, None,
                        st.grad(m.norm_w), st.grad(m.norm_b), dx0_out=dseq[:, nF:])
        _done(m)
        return dseq, None, None, None, None


# --------------------------------------------------------------------------------------------
# a11  classifier tail: tokens.mean(1) -> BatchNorm1d(affine=False) -> Linear      classifier.py:49-58
# --------------------------------------------------------------------------------------------
class ClassifierTailFn(torch.autograd.Function):
    """All f32 (the reference runs this path without autocast).  ``anchor`` (the head weight) makes autograd
    record the node when the encoder is frozen and ``x`` carries no gradient."""

    @staticmethod
    def forward(ctx, x: Tensor, anchor: Tensor, m: SimpleNamespace, bn_training: bool):
        x = x.contiguous()
        B, n, D = x.shape
        pooled = K.meanpool_fwd(x)
        mean = rstd = None
        feat = pooled
        if m.bn is not None:
            bn = m.bn
            momentum = 0.1 if bn.momentum is None else bn.momentum
            feat, mean, rstd = K.batchnorm1d_fwd(pooled, bn.running_mean, bn.running_var, bn_training, momentum, bn.eps)
            if bn_training and bn.num_batches_tracked is not None:
                bn.num_batches_tracked += 1
        y = K.head_fwd(feat, m.head_w.data, m.head_b.data)
        ctx.m, ctx.n, ctx.bn_training = m, n, bn_training
        ctx.save_for_backward(pooled, feat, mean, rstd)
        return y

    @staticmethod
    def backward(ctx, dy: Tensor):
        m = ctx.m
        st: ParamStore = m.store
        pooled, feat, mean, rstd = ctx.saved_tensors
        dy = dy.contiguous().float()
        need_dx = ctx.needs_input_grad[0]
        dfeat = K.head_bwd(dy, feat, m.head_w.data,
                           st.grad(m.head_w) if m.head_w.requires_grad else None,
                           st.grad(m.head_b) if m.head_b.requires_grad else None, need_dx)
        _done(m)
        if not need_dx:
            return None, None, None, None
        dpool = K.batchnorm1d_bwd(dfeat, pooled, mean, rstd, ctx.bn_training) if m.bn is not None else dfeat
        return K.meanpool_bwd(dpool, ctx.n), None, None, None

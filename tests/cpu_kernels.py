"""Torch (CPU or CUDA) emulation of every wrapper in deepavfusion_b200/kernels.py.

TEST INFRASTRUCTURE ONLY.  Two uses:
  * ``-m "not gpu"`` tests monkey-patch ``deepavfusion_b200.kernels`` with these functions to check
    the host-side orchestration (autograd wiring, shapes, state_dict, gradient routing) against the
    oracle on CPU;
  * ``-m gpu`` tests call the real CUDA kernel and this emulation on the same inputs, op by op.
Each function states the same contract as the kernel it mirrors (bf16 rounding points included).
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn.functional as F

ACT_NONE, ACT_GELU, ACT_DGELU = 0, 1, 2
bf16 = torch.bfloat16


def mask_rank(noise, len_keep):
    B, L = noise.shape
    ids_shuffle = torch.argsort(noise, dim=1, stable=True)
    ids_restore = torch.argsort(ids_shuffle, dim=1, stable=True)
    ids_keep = ids_shuffle[:, :len_keep].contiguous()
    mask = (ids_restore >= len_keep).float()
    return ids_restore, ids_keep, mask


def patch_rows(img, ids_keep, p):
    B, C, H, W = img.shape
    gH, gW = H // p, W // p
    rows = img.reshape(B, C, gH, p, gW, p).permute(0, 2, 4, 1, 3, 5).reshape(B, gH * gW, C * p * p)
    if ids_keep is not None:
        rows = rows.gather(1, ids_keep.unsqueeze(-1).expand(-1, -1, rows.shape[-1]))
    return rows.reshape(-1, C * p * p).to(bf16)


def _rowmap(M, g, G, off, device):
    m = torch.arange(M, device=device)
    return (m // g) * G + off + (m % g)


def cast_rows_bf16(src, M=None, g=None, G=None, off=0):
    D = src.shape[-1]
    s2 = src.reshape(-1, D)
    if M is None:
        return s2.to(bf16)
    return s2[_rowmap(M, g, G, off, src.device)].to(bf16)


def cast_flat_bf16(src, dst):
    dst.copy_(src.to(bf16))


def sum_cast(parts, want_bf16=True):
    out = parts[0] + parts[1]
    if len(parts) > 2:
        out = out + parts[2]
    return out, (out.to(bf16) if want_bf16 else None)


def colsum_bf16(x, out):
    out += x.float().sum(0)


def batchsum_f32(x, off, g, out, accumulate):
    s = x[:, off:off + g].sum(0)
    if accumulate:
        out += s
    else:
        out.copy_(s)


def _seg_perm(B, n, seg_start, device):
    """index such that seg_major[k] = natural[perm[k]] for the segmented bf16 row layout."""
    if seg_start is None or len(seg_start) <= 2:
        return None
    perm = []
    for k in range(len(seg_start) - 1):
        st, en = seg_start[k], seg_start[k + 1]
        b = torch.arange(B, device=device).view(B, 1) * n
        r = torch.arange(st, en, device=device).view(1, -1)
        perm.append((b + r).reshape(-1))
    return torch.cat(perm)


def layernorm_fwd(x0, x1, gamma, beta, eps, want_bf16=True, want_f32=False, seg_start=None):
    x = x0 if x1 is None else torch.cat([x0, x1], dim=1)
    B, n, D = x.shape
    x = x.reshape(B * n, D).float()
    mean = x.mean(-1)
    var = x.var(-1, unbiased=False)
    rstd = torch.rsqrt(var + eps)
    y = (x - mean[:, None]) * rstd[:, None] * gamma + beta
    yb = None
    if want_bf16:
        perm = _seg_perm(B, n, seg_start, x.device)
        yb = (y if perm is None else y[perm]).to(bf16).contiguous()
    return yb, (y.contiguous() if want_f32 else None), mean, rstd


def layernorm_bwd(x0, x1, gamma, mean, rstd, dy_bf16, dy_f32, add0, add1, dgamma, dbeta, seg_start=None,
                  need_dx0=True, need_dx1=True, dx0_out=None, dx0_lowp=None, dx1_lowp=None):
    x = x0 if x1 is None else torch.cat([x0, x1], dim=1)
    B, n, D = x.shape
    n0 = x0.shape[1]
    x = x.reshape(B * n, D).float()
    dy = torch.zeros_like(x)
    if dy_bf16 is not None:
        perm = _seg_perm(B, n, seg_start, x.device)
        d = dy_bf16.float().reshape(B * n, D)
        if perm is None:
            dy += d
        else:
            dy[perm] += d
    if dy_f32 is not None:
        dy += dy_f32.reshape(B * n, D)
    xh = (x - mean[:, None]) * rstd[:, None]
    dgamma += (dy * xh).sum(0)
    dbeta += dy.sum(0)
    g = dy * gamma
    dx = rstd[:, None] * (g - g.mean(-1, keepdim=True) - xh * (g * xh).mean(-1, keepdim=True))
    dx = dx.reshape(B, n, D)
    dx0 = dx[:, :n0]
    if add0 is not None:
        dx0 = dx0 + add0.reshape(B, n0, D)
    dx1 = None
    if x1 is not None:
        dx1 = dx[:, n0:]
        if add1 is not None:
            dx1 = dx1 + add1.reshape(B, n - n0, D)
        dx1 = dx1.contiguous() if need_dx1 else None
    if dx0_lowp is not None:
        dx0_lowp.copy_(dx0.reshape(dx0_lowp.shape))
    if dx1_lowp is not None:
        dx1_lowp.copy_(dx1.reshape(dx1_lowp.shape))
    if dx0_out is not None:
        dx0_out.copy_(dx0)
        dx0 = dx0_out
    else:
        dx0 = dx0.contiguous() if need_dx0 else None
    return dx0, dx1


def _gelu(x):
    return 0.5 * x * (1 + torch.erf(x * 0.7071067811865476))


def _dgelu(x):
    return 0.5 * (1 + torch.erf(x * 0.7071067811865476)) + x * torch.exp(-0.5 * x * x) * 0.3989422804014327


def gemm(a, b, a_kmajor=True, b_kmajor=True, *, bias=None, act=ACT_NONE, want_aux=False, aux_in=None, res=None,
         res_idx=None, out=None, out_dtype=torch.bfloat16, accumulate=False, window=None, split_k=0, rowsum_out=None, debug_clocks=None):
    A = a.float() if a_kmajor else a.float().t()
    if rowsum_out is not None:
        rowsum_out += A.sum(1)
    Bm = b.float() if b_kmajor else b.float().t()
    z = A @ Bm.t()
    M, N = z.shape
    if bias is not None:
        z = z + bias
    aux = (_dgelu(z) if act == ACT_GELU else z).to(bf16) if want_aux else None     # next to GELU the aux output is gelu'(z)
    if act == ACT_GELU:
        z = _gelu(z)
    elif act == ACT_DGELU:
        z = z * aux_in.float().reshape(M, -1)[:, :N]
    rows = torch.arange(M, device=z.device) if window is None else _rowmap(M, *window, z.device)
    if res is not None:
        r2 = res.reshape(-1, res.shape[-1])
        z = z + r2[res_idx if res_idx is not None else rows][:, :N]
    if out is None:
        out = torch.empty(M, N, dtype=out_dtype, device=z.device)
    o2 = out.reshape(-1, out.shape[-1]) if out.is_contiguous() else out
    if accumulate:
        o2[rows, :N] = o2[rows, :N] + z
    else:
        o2[rows, :N] = z.to(out.dtype)
    return (out, aux) if want_aux else out


def gemm_grouped(problems):
    return [gemm(*pos, **kw) for pos, kw in problems]


def attention_fwd(q, k, v, scale, out=None, accumulate=False):
    qf, kf, vf = q.float().permute(0, 2, 1, 3), k.float().permute(0, 2, 1, 3), v.float().permute(0, 2, 1, 3)
    s = qf @ kf.transpose(-1, -2) * scale
    lse = torch.logsumexp(s, dim=-1)
    o = (torch.softmax(s, -1) @ vf).permute(0, 2, 1, 3)
    if out is None:
        out = torch.empty(o.shape, dtype=bf16, device=q.device)
    if accumulate:
        out.copy_((out.float() + o).to(bf16))
    else:
        out.copy_(o.to(bf16))
    return out, lse.contiguous()


def attention_bwd(q, k, v, d_o, lse, scale, dq, dk, dv, accumulate_dq=False, o=None, dq_dead_rows=0):
    if dq_dead_rows:          # dq is a view [B, n, H, d] of a larger buffer: zero the rows in front of it
        B_, n_, H_, d_ = dq.shape
        torch.as_strided(dq, (B_, dq_dead_rows, H_, d_), dq.stride(), dq.storage_offset() - dq_dead_rows * dq.stride(1)).zero_()
    qf, kf, vf = q.float().permute(0, 2, 1, 3), k.float().permute(0, 2, 1, 3), v.float().permute(0, 2, 1, 3)
    dof = d_o.float().permute(0, 2, 1, 3)
    p = torch.exp(qf @ kf.transpose(-1, -2) * scale - lse[..., None])
    dp = dof @ vf.transpose(-1, -2)
    Dl = (p * dp).sum(-1, keepdim=True) if o is None else (dof * o.float().permute(0, 2, 1, 3)).sum(-1, keepdim=True)
    ds = p * (dp - Dl) * scale
    dqv = (ds @ kf).permute(0, 2, 1, 3)
    if accumulate_dq:
        dqv = dqv + dq.float()
    dq.copy_(dqv.to(bf16))
    dk.copy_((ds.transpose(-1, -2) @ qf).permute(0, 2, 1, 3).to(bf16))
    dv.copy_((p.transpose(-1, -2) @ dof).permute(0, 2, 1, 3).to(bf16))


def decoder_assemble_fwd(e, ef, mask_token, pos, ids_restore, nK, nF):
    B, L = ids_restore.shape
    D = e.shape[-1]
    x = torch.cat([e.reshape(B, nK, D), mask_token.reshape(1, 1, D).expand(B, L - nK, D)], 1)
    x = x.gather(1, ids_restore.unsqueeze(-1).expand(-1, -1, D)) + pos.reshape(1, L, D)
    return torch.cat([ef.reshape(B, nF, D), x], 1).contiguous()


def decoder_assemble_bwd(dseq, ids_keep, ids_restore, nF, dmask_token, dpos):
    B, S, D = dseq.shape
    L, nK = S - nF, ids_keep.shape[1]
    dx = dseq[:, nF:]
    dpos += dx.sum(0)
    masked = (ids_restore >= nK).float().unsqueeze(-1)
    dmask_token += (dx * masked).sum((0, 1))
    de = dx.gather(1, ids_keep.unsqueeze(-1).expand(-1, -1, D)).reshape(B * nK, D).to(bf16)
    df = dseq[:, :nF].reshape(B * nF, D).to(bf16)
    return de, df


def _targets(img, p, norm_pix):
    B, C, H, W = img.shape
    gH, gW = H // p, W // p
    t = img.reshape(B, C, gH, p, gW, p)
    t = torch.einsum("nchpwq->nhwpqc", t).reshape(B, gH * gW, p * p * C)
    if norm_pix:
        t = (t - t.mean(-1, keepdim=True)) / (t.var(-1, keepdim=True) + 1e-6) ** 0.5
    return t


def masked_mse_fwd(img, pred, mask, p, pred_G, pred_off, norm_pix):
    B = img.shape[0]
    t = _targets(img, p, norm_pix)
    L, P = t.shape[1], t.shape[2]
    pr = pred.reshape(B, pred_G, P)[:, pred_off:pred_off + L]
    return (((pr - t) ** 2).mean(-1) * mask).sum().reshape(1)


def masked_mse_bwd(img, pred, mask, gscale, inv_count, p, pred_G, pred_off, norm_pix):
    B = img.shape[0]
    t = _targets(img, p, norm_pix)
    L, P = t.shape[1], t.shape[2]
    pr = pred.reshape(B, pred_G, P)[:, pred_off:pred_off + L]
    d = gscale.reshape(()) * inv_count * 2.0 / P * (pr - t) * mask.unsqueeze(-1)
    return d.reshape(B * L, P).to(bf16)


def adamw_step(p, g, m, v, p_bf16, chunk_group, hp, scal, beta1, beta2, eps, zero_grad, sumsq_out=None):
    bc1, bc2, gsc = 1 - float(scal[0]), 1 - float(scal[1]), float(scal[2])
    gid = chunk_group.long().repeat_interleave(64)
    live = gid != 255
    gsafe = gid.clamp(max=hp.numel() // 2 - 1)
    lr, wd = hp[2 * gsafe], hp[2 * gsafe + 1]
    gr = g * gsc
    if sumsq_out is not None:
        sumsq_out += (gr[live].double() ** 2).sum().float()
    m_new = beta1 * m + (1 - beta1) * gr
    v_new = beta2 * v + (1 - beta2) * gr * gr
    denom = v_new.sqrt() / math.sqrt(bc2) + eps
    p_new = p * (1 - lr * wd) - (lr / bc1) * (m_new / denom)
    m.copy_(torch.where(live, m_new, m)); v.copy_(torch.where(live, v_new, v)); p.copy_(torch.where(live, p_new, p))
    if zero_grad:
        g.zero_()
    if p_bf16 is not None:
        p_bf16.copy_(torch.where(live, p.to(bf16), p_bf16))


def sumsq_f32(g, out):
    out += (g.double() ** 2).sum().float()


def meanpool_fwd(x):
    return x.float().mean(dim=1)


def meanpool_bwd(dy, n):
    return (dy / n)[:, None, :].expand(-1, n, -1).contiguous()


def batchnorm1d_fwd(x, running_mean, running_var, training, momentum, eps):
    B = x.shape[0]
    if training:
        mean, var = x.mean(0), x.var(0, unbiased=False)
        running_mean.mul_(1 - momentum).add_(momentum * mean)
        running_var.mul_(1 - momentum).add_(momentum * (x.var(0, unbiased=True) if B > 1 else var))
    else:
        mean, var = running_mean.clone(), running_var.clone()
    rstd = (var + eps).rsqrt()
    return (x - mean) * rstd, mean, rstd


def batchnorm1d_bwd(dy, x, mean, rstd, training):
    if not training:
        return dy * rstd
    xh = (x - mean) * rstd
    return rstd * (dy - dy.mean(0) - xh * (dy * xh).mean(0))


def head_fwd(x, W, bias):
    y = x @ W.t()
    return y + bias if bias is not None else y


def head_bwd(dy, x, W, dW, db, need_dx):
    if dW is not None:
        dW += dy.t() @ x
    if db is not None:
        db += dy.sum(0)
    return dy @ W if need_dx else None


def scale_rows_add(res, y, scale, rows_per_sample):
    D = res.shape[-1]
    s = scale.repeat_interleave(rows_per_sample)[:, None]
    return (res.reshape(-1, D) + s * y.reshape(-1, D)).reshape(res.shape)


def scale_rows(src, scale, rows_per_sample, want_f32=False, want_bf16=True):
    D = src.shape[-1]
    v = scale.repeat_interleave(rows_per_sample)[:, None] * src.reshape(-1, D)
    return (v if want_f32 else None), (v.to(bf16) if want_bf16 else None)


def launch_count():
    return 0


def set_gemm_impl(impl):
    pass


def set_attn_impl(impl):
    pass


def set_gemm_2cta(on):
    pass


def set_gemm_sms(n):
    return n


def install(monkeypatch=None):
    """Replace every kernel wrapper in deepavfusion_b200.kernels by its emulation."""
    import deepavfusion_b200.kernels as K
    import deepavfusion_b200.models.layers as Lyr
    names = [n for n, f in globals().items() if callable(f) and not n.startswith("_") and n not in ("install", "Optional")]
    for n in names:
        if hasattr(K, n):
            if monkeypatch is not None:
                monkeypatch.setattr(K, n, globals()[n])
            else:
                setattr(K, n, globals()[n])
    if monkeypatch is not None:
        monkeypatch.setattr(Lyr, "_REQUIRE_CUDA", False)
    else:
        Lyr._REQUIRE_CUDA = False

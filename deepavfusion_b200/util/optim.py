"""Fused AdamW (K13/K14): a ``torch.optim.Optimizer`` whose ``step()`` is ONE kernel launch over the
flat parameter / gradient / moment buffers of a ``ParamStore``.

Replaces ``torch.optim.AdamW(param_groups, lr, betas=(0.9, 0.95))`` (train.py:93) together with the
per-step sweeps around it in ``Trainer.step`` (misc.py:109-134): gradient / accum_iter, global
gradient norm, ``zero_grad`` and the bf16 weight casts.  ``param_groups`` keep torch's layout so
``util.lr_sched.adjust_learning_rate`` (lr_sched.py:18-23) and ``state_dict()`` consumers work
unchanged; ``state[p]['exp_avg' | 'exp_avg_sq']`` are views into the flat moment buffers and
``state[p]['step']`` is torch's per-parameter step tensor, so an ``'optimizer'`` checkpoint entry
is loadable by ``torch.optim.AdamW`` and vice versa.
"""
from __future__ import annotations

from typing import Optional

import torch

from .. import kernels as K
from ..params import ALIGN, ParamStore


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, store: ParamStore, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        defaults = dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        self.store = store
        dev = store.flat_p.device
        self.flat_m = torch.zeros_like(store.flat_p)
        self.flat_v = torch.zeros_like(store.flat_p)
        ngroups = len(self.param_groups)
        assert ngroups < 255, "at most 254 parameter groups"
        cg = torch.full((store.numel // ALIGN,), 255, dtype=torch.uint8)
        for gi, group in enumerate(self.param_groups):
            assert tuple(group["betas"]) == tuple(self.param_groups[0]["betas"]) and group["eps"] == self.param_groups[0]["eps"], \
                "per-group betas / eps are not supported by the fused kernel"
            for p in group["params"]:
                k = store.index_of(p)
                b, e = store.span(k)
                if p.requires_grad:
                    # the stacked pair-attention k / v tensors share a 64-element chunk when k ends off the ALIGN grid
                    # (ParamStore.stacked): a chunk has ONE group, so both neighbours must agree on it
                    c0, c1 = b // ALIGN, (b + p.numel() + ALIGN - 1) // ALIGN
                    taken = cg[c0:c1]
                    assert bool(((taken == 255) | (taken == gi)).all()), \
                        f"{store.names[k]}: shares a {ALIGN}-element chunk with a parameter of another group"
                    cg[c0:c1] = gi
                self.state[p] = dict(step=torch.zeros((), dtype=torch.float32),
                                     exp_avg=self.flat_m[b:b + p.numel()].view(p.shape),
                                     exp_avg_sq=self.flat_v[b:b + p.numel()].view(p.shape))
        self.chunk_group = cg.to(dev)
        self.hp = torch.zeros(2 * ngroups, dtype=torch.float32, device=dev)
        # ring of pinned staging buffers: the host runs several (graph-replayed) steps ahead of the GPU, so the buffer of
        # step t must not be rewritten before its queued H2D copy has executed; slot i is reused only after its event
        self._hp_ring = [torch.zeros(2 * ngroups, dtype=torch.float32, pin_memory=dev.type == "cuda") for _ in range(8)]
        self._hp_events = [None] * len(self._hp_ring)
        self._hp_slot = 0
        self._hp_last = None
        # {beta1^t, beta2^t, grad_scale, -}: advanced on the device so a captured step replays correctly
        self.scal = torch.tensor([1.0, 1.0, 1.0, 0.0], dtype=torch.float32, device=dev)
        b1, b2 = self.param_groups[0]["betas"]
        self._beta_mul = torch.tensor([b1, b2, 1.0, 1.0], dtype=torch.float32, device=dev)
        self.grad_sumsq = torch.zeros(1, dtype=torch.float32, device=dev)
        self.n_steps = 0
        self._open = False

    # -- hyper-parameter upload (only when the schedule changed something) -----------------------
    def _sync_hp(self):
        vals = []
        for g in self.param_groups:
            vals += [float(g["lr"]), float(g["weight_decay"])]
        if vals != self._hp_last:
            k = self._hp_slot
            self._hp_slot = (k + 1) % len(self._hp_ring)
            if self._hp_events[k] is not None:
                self._hp_events[k].synchronize()            # the copy that last read this slot (8 uploads ago) is done
            host = self._hp_ring[k]
            host.copy_(torch.tensor(vals, dtype=torch.float32))
            self.hp.copy_(host, non_blocking=True)
            if self.hp.is_cuda:
                ev = torch.cuda.Event()
                ev.record()
                self._hp_events[k] = ev
            self._hp_last = vals

    def set_grad_scale(self, scale: float) -> None:
        """1 / (accum_iter * world_size): folded into the update instead of a separate sweep."""
        self.scal[2] = scale

    @torch.no_grad()
    def begin_step(self, sync_hp: bool = True):
        """Opens an optimizer step that is applied bucket by bucket while backward still runs (``step_range``,
        driven by ``GradSync``): advances beta^t and clears the grad-norm accumulator once."""
        if sync_hp:
            self._sync_hp()
        self.scal.mul_(self._beta_mul)                      # beta^t on the device
        self.grad_sumsq.zero_()
        self._open = True

    @torch.no_grad()
    def step_range(self, lo: int, hi: int, zero_grad: bool = True):
        """The fused update on elements [lo, hi) of the flat buffers (a gradient bucket whose values are final)."""
        assert self._open and lo % ALIGN == 0 and hi % ALIGN == 0
        b1, b2 = self.param_groups[0]["betas"]
        st = self.store
        K.adamw_step(st.flat_p[lo:hi], st.flat_g[lo:hi], self.flat_m[lo:hi], self.flat_v[lo:hi], st.flat_lp[lo:hi],
                     self.chunk_group[lo // ALIGN:hi // ALIGN], self.hp, self.scal,
                     float(b1), float(b2), float(self.param_groups[0]["eps"]), zero_grad, self.grad_sumsq)

    @torch.no_grad()
    def end_step(self):
        """Closes a bucketed step (every range has been launched)."""
        assert self._open
        self._open = False
        self.store.mark_lowp_fresh()
        self.n_steps += 1

    @torch.no_grad()
    def step(self, closure=None, zero_grad: bool = True, sync_hp: bool = True):
        """One fused launch: p, m, v update (+bf16 shadow, +grad-norm^2, +zero_grad)."""
        assert closure is None and not self._open
        self.begin_step(sync_hp)
        self.step_range(0, self.store.numel, zero_grad)
        self.end_step()
        return None

    def grad_norm(self) -> torch.Tensor:
        """Global L2 norm of the (scaled) gradients consumed by the last step(); device tensor, no sync."""
        return self.grad_sumsq.sqrt().reshape(())

    def zero_grad(self, set_to_none: bool = False):
        self.store.zero_grad()

    # -- checkpoint layout == torch.optim.AdamW ----------------------------------------------------
    def state_dict(self):
        for st_ in self.state.values():
            st_["step"].fill_(float(self.n_steps))
        return super().state_dict()

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)                 # replaces the state tensors by loaded copies
        steps = 0
        for group in self.param_groups:
            for p in group["params"]:
                s = self.state[p]
                k = self.store.index_of(p)
                b, _ = self.store.span(k)
                m_view = self.flat_m[b:b + p.numel()].view(p.shape)
                v_view = self.flat_v[b:b + p.numel()].view(p.shape)
                m_view.copy_(s["exp_avg"]); v_view.copy_(s["exp_avg_sq"])
                steps = max(steps, int(float(s["step"])))
                s["exp_avg"], s["exp_avg_sq"] = m_view, v_view
                s["step"] = torch.tensor(float(steps), dtype=torch.float32)
        self.n_steps = steps
        b1, b2 = self.param_groups[0]["betas"]
        self.scal[0] = b1 ** steps
        self.scal[1] = b2 ** steps
        self._hp_last = None

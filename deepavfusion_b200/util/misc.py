"""Training-step runtime (mirror of reference util/misc.py:27-163 ``Trainer``).

Same constructor and methods (``autocast``, ``autosync``, ``step``, ``backward``, ``zero_grad``,
``get_scale``, ``module_dict``) so ``train.py:97-103,163-170`` runs unchanged.  What differs is what
they call: backward is the hand-written CUDA path, the DDP reducer is ``GradSync`` (bucketed NCCL
all-reduce overlapped with backward), and unscale + /accum + grad-norm + AdamW + zero_grad + bf16
weight refresh are ONE fused kernel.  The two per-step host syncs of the reference
(train.py:166, misc.py:78-79) are gone: ``step`` returns the gradient norm as a 0-dim DEVICE tensor
(``MetricLogger.update`` calls ``.item()`` on tensors, meters.py:101-102).
"""
from __future__ import annotations

import contextlib
import os
import math

import torch

from ..models.layers import ensure_store
from . import distributed as dist_utils
from .optim import FusedAdamW


class Trainer:
    def __init__(self, model, criterion=None, optimizer=None, accum_iter=1, use_amp=True, distributed=False,
                 bucket_mb: float = 64.0):
        """``optimizer``: a ``FusedAdamW``, or a ``torch.optim.AdamW`` built the reference way
        (train.py:89-93) whose param groups / hyper-parameters are adopted by a FusedAdamW.
        ``use_amp`` is accepted for signature compatibility: the tensor-core path always computes in bf16
        with f32 accumulation (BASELINE.json), which needs no GradScaler."""
        self.distributed = bool(distributed) and dist_utils.get_world_size() > 1
        self.model_without_ddp = model
        self.model = model
        self.n_steps = torch.tensor([0])
        self.criterion = criterion
        self.store = ensure_store(model)
        if optimizer is not None and not isinstance(optimizer, FusedAdamW):
            if not isinstance(optimizer, torch.optim.AdamW):
                raise TypeError("Trainer drives the fused AdamW kernel; pass a torch.optim.AdamW (train.py:93) or a FusedAdamW")
            groups = [{k: v for k, v in g.items() if k in ("params", "lr", "betas", "eps", "weight_decay", "lr_scale", "pretrained")}
                      for g in optimizer.param_groups]
            d = optimizer.defaults
            optimizer = FusedAdamW(groups, self.store, lr=d["lr"], betas=d["betas"], eps=d["eps"], weight_decay=d["weight_decay"])
        self.optimizer = optimizer
        self.scaler = None
        self.accum_iter = accum_iter
        self.accums = 0
        # bucket scheduler: gradient all-reduce (N > 1) and / or the optimizer step overlapped with backward
        mode = os.environ.get("DAVF_OVERLAP_ADAMW", "1")         # "0": off; "force": also on CPU tensors (host-logic tests)
        overlap_opt = isinstance(self.optimizer, FusedAdamW) and mode != "0" and (self.store.flat_g.is_cuda or mode == "force")
        self.sync = dist_utils.GradSync(self.store, bucket_mb=bucket_mb, optimizer=self.optimizer if overlap_opt else None) \
            if (self.distributed or overlap_opt) else None
        self.grad_buffer_registered = "single process"
        if self.distributed:
            if self.store.flat_g.is_cuda:
                # the gradient buffer lives in NCCL-registered memory: NVLS all-reduces run in place on it
                buf, pool = dist_utils.nccl_registered_zeros(self.store.numel, self.store.device)
                if buf is not None:
                    self.store.rehome_grads(buf)
                    self._nccl_pool, self.grad_buffer_registered = pool, True
                else:
                    self.grad_buffer_registered = pool          # the reason
            self.broadcast_parameters()
            # SMs left to the NCCL kernels that run under backward (see davf_set_gemm_sms).  Only BACKWARD launches of the
            # final micro-step overlap the bucket all-reduces: forward (and accumulate-only) launches keep the whole machine.
            self._comm_sms = int(os.environ.get("DAVF_COMM_SMS", "32")) if self.store.flat_g.is_cuda else 0
        world = dist_utils.get_world_size() if self.distributed else 1
        if self.optimizer is not None:
            self.optimizer.set_grad_scale(1.0 / (self.accum_iter * world))
        self.eval_model = self.model_without_ddp
        self.zero_grad()

    def broadcast_parameters(self):
        """Rank 0 -> all, once (what the DDP constructor does, misc.py:34): one flat buffer, then the module buffers.
        The reference converts BatchNorm to SyncBatchNorm first (misc.py:33); the classifier's BatchNorm1d kernel
        (AVClassifier(input_norm=True)) normalises with per-rank batch statistics, so a data-parallel run of it would
        silently diverge from the reference: refuse it."""
        for mod in self.model_without_ddp.modules():
            if isinstance(mod, torch.nn.modules.batchnorm._BatchNorm):
                raise NotImplementedError("data-parallel training with BatchNorm (AVClassifier(input_norm=True)) needs "
                                          "SyncBatchNorm statistics (misc.py:33), which the classifier tail kernel does not "
                                          "all-reduce; run lin-probe on one GPU or with input_norm=False")
        torch.distributed.broadcast(self.store.flat_p, src=0)
        for b in self.model_without_ddp.buffers():
            torch.distributed.broadcast(b, src=0)
        self.store.refresh_lowp(force=True)

    def module_dict(self):
        d = {"state_dict": self.model_without_ddp, "n_steps": self.n_steps}
        if self.criterion is not None:
            d["criterion"] = self.criterion
        if self.optimizer is not None:
            d["optimizer"] = self.optimizer
        return d

    def zero_grad(self):
        self.store.zero_grad()
        self.accums = 0

    def get_scale(self):
        return 1.0

    def backward(self, loss, create_graph=False, _fuse_optimizer=False, _sync_hp=True):
        assert not create_graph
        if self.sync is not None:
            self.sync.enabled = self.accums == self.accum_iter - 1       # all-reduce on the last micro-step only
            if _fuse_optimizer and self.sync.enabled and self.sync.optimizer is not None:
                self.optimizer.begin_step(sync_hp=_sync_hp)              # the step is applied bucket by bucket during backward
                self.sync.fuse_optimizer = True
        comm_sms = getattr(self, "_comm_sms", 0) if (self.sync is not None and self.sync.enabled and self.distributed) else 0
        if comm_sms:
            from .. import kernels as K
            K.set_gemm_sms(148 - comm_sms)
        try:
            loss.backward()
        finally:
            if comm_sms:
                K.set_gemm_sms(148)
        self.store.join_side_streams(torch.cuda.current_stream() if self.store.flat_g.is_cuda else None)
        if self.sync is not None:
            self.sync.finish()
        self.accums += 1

    def step(self, loss, create_graph=False, clip_grad=None, skip_grad=None):
        if clip_grad is not None or skip_grad is not None:
            raise NotImplementedError("clip_grad / skip_grad are unset in every pre-training config (deepavfusion.yaml:60)")
        return self._step(loss)

    def _step(self, loss, sync_hp=True):
        fused = self.sync is not None and self.sync.optimizer is not None
        self.backward(loss, _fuse_optimizer=fused, _sync_hp=sync_hp)
        norm = None
        if self.accums == self.accum_iter:
            if not fused:
                self.optimizer.step(zero_grad=True, sync_hp=sync_hp)   # /accum/world, grad-norm^2, AdamW, bf16 refresh, zero_grad
            norm = self.optimizer.grad_norm()
            self.accums = 0
            self.n_steps += 1
        return norm, 1.0

    def autocast(self):
        return contextlib.nullcontext()

    def autosync(self):
        return contextlib.nullcontext()


def get_grad_norm_(parameters, norm_type: float = 2.0) -> torch.Tensor:
    """misc.py:151-163 for callers that still want it (one kernel over the flat buffer when possible)."""
    parameters = [p for p in ([parameters] if isinstance(parameters, torch.Tensor) else parameters) if p.grad is not None]
    if not parameters:
        return torch.tensor(0.0)
    if norm_type == math.inf:
        return max(p.grad.detach().abs().max() for p in parameters)
    return torch.norm(torch.stack([torch.norm(p.grad.detach(), norm_type) for p in parameters]), norm_type)


class CheckpointManager:
    """Mirror of reference util/misc.py:222-309: same constructor, ``resume()`` / ``checkpoint(epoch, save_dict, is_best)``
    and the same file layout -- ``checkpoint_latest.pth`` (plus ``checkpoint_best.pth`` / ``checkpoint_{epoch:04d}.pth``)
    holding ``{name: module.state_dict() or tensor}`` for every entry of ``Trainer.module_dict()`` ('state_dict',
    'optimizer', 'n_steps', ...) together with 'epoch' and the caller's metrics.  Because the fused optimizer's
    ``state_dict()`` has ``torch.optim.AdamW``'s shape, files written here resume under the reference and vice versa.
    Tensors are saved as independent CPU copies (the parameters are views into one flat buffer)."""

    def __init__(self, modules, ckpt_dir, epochs, save_freq=None):
        import os
        self.modules, self.ckpt_dir, self.epochs, self.save_freq = modules, ckpt_dir, epochs, save_freq
        self.world_size, self.rank = dist_utils.get_world_size(), dist_utils.get_rank()
        if self.rank == 0:
            os.makedirs(self.ckpt_dir, exist_ok=True)

    @staticmethod
    def _to_cpu(state):
        if isinstance(state, dict):
            return {k: CheckpointManager._to_cpu(v) for k, v in state.items()}
        if isinstance(state, (list, tuple)):
            return type(state)(CheckpointManager._to_cpu(v) for v in state)
        if isinstance(state, torch.Tensor):
            return state.detach().to("cpu", copy=True).contiguous()
        return state

    def create_state_dict(self, save_dict=None):
        state = {}
        for k, mod in self.modules.items():
            if mod is None:
                state[k] = None
            elif isinstance(mod, torch.Tensor):
                state[k] = mod.detach().to("cpu", copy=True)
            else:
                state[k] = self._to_cpu(mod.state_dict())
        if save_dict is not None:
            state.update(save_dict)
        return state

    def resume(self):
        import os
        fname = os.path.join(self.ckpt_dir, "checkpoint_latest.pth")
        start_epoch, metrics = 0, {}
        if os.path.isfile(fname):
            ckpt = torch.load(fname, map_location="cpu", weights_only=False)
            for k, mod in self.modules.items():
                if mod is None:
                    continue
                if isinstance(mod, torch.Tensor):
                    mod.data[:] = ckpt[k].data
                else:
                    mod.load_state_dict(ckpt[k])
            start_epoch = ckpt["epoch"]
            metrics = {k: v for k, v in ckpt.items() if k not in self.modules and k != "epoch"}
            print(f"=> loaded checkpoint '{fname}' (epoch {start_epoch})")
        return start_epoch, metrics

    def checkpoint(self, epoch, save_dict=None, is_best=False):
        import os
        if self.rank != 0:
            return
        state = self.create_state_dict(save_dict)
        state.setdefault("epoch", epoch)
        targets = ["checkpoint_latest.pth"]
        if is_best:
            targets.append("checkpoint_best.pth")
        if self.save_freq is not None and (epoch % self.save_freq == 0 or epoch == self.epochs):
            targets.append(f"checkpoint_{epoch:04d}.pth")
        for t in targets:
            torch.save(state, os.path.join(self.ckpt_dir, t))

"""Whole-step CUDA graph: forward + backward (+ bucket all-reduce) + fused AdamW captured once and
replayed, removing the ~2,000 per-step launches' CPU cost and every host sync from the loop
(eager PyTorch on this path is launch-bound).  Hyper-parameters (lr / weight decay,
beta^t, grad scale) live in device tables the captured kernels read, so LR schedules need no
re-capture; the mask noise comes from torch's graph-safe Philox generator state.

Gradient accumulation (``Trainer.accum_iter > 1``: the AudioSet pre-training and fine-tuning recipes,
README.md:51-55, configs/finetune.yaml:36-50) uses TWO graphs, as misc.py:144-148 has two kinds of micro-step:
``graph_micro`` = forward + backward accumulating into the flat gradient buffer (DDP ``no_sync``: no all-reduce, no
optimizer), replayed on the first accum_iter - 1 micro-steps, and ``graph_final`` = forward + backward + bucketed
all-reduce + fused AdamW (which also divides by accum_iter * world and zeroes the gradients) on the last one.

Data-parallel runs capture the bucketed gradient all-reduce too (``overlap_comm``, the default): every bucket's
NCCL call is recorded on the communication stream at the point of backward where its last gradient has been
produced, so in the replayed graph the all-reduce of the decoder / late-encoder buckets runs under the rest of
backward, and each bucket's fused AdamW range launch follows its all-reduce on a third stream.  NCCL needs its communicator and internal stream
to exist before capture: the eager warm-up steps run the identical bucketed path.  ``DAVF_GRAPH_NCCL=0`` (or
``overlap_comm=False``) falls back to capturing forward + backward only, with one all-reduce over the flat
gradient buffer and the AdamW launch issued after the replay.

Side effects of construction.  The eager warm-up steps and the capture pass exist to instantiate kernels, TMA
descriptors, allocator pools and the NCCL communicator; they must not train.  Everything they touch -- parameters,
bf16 shadows, gradients, Adam moments, beta^t, the optimizer / trainer step counters, module buffers (BatchNorm
running statistics) -- is snapshotted before and restored after, so the first replay starts from exactly the state
the caller handed in.  What is NOT restored: the CUDA RNG offset (the warm-up draws mask noise; seed after
constructing this object if a run must be reproducible draw for draw)."""
from __future__ import annotations

import os
from typing import Callable, Optional, Sequence

import torch


def _avmae_loss(model, image, audio):
    """Default step body: train.py:163-165 (loss = loss_image + loss_audio; both are also the logged metrics)."""
    li, la, _, _ = model(image, audio)
    return li + la, (li.detach(), la.detach())


class GraphedTrainStep:
    """Capture once, replay per step.  Drop references to the outputs (losses, predictions) of earlier EAGER
    steps before constructing this object: they carry ``record_stream`` marks from the multi-stream forward, and
    freeing such a tensor while a capture is in progress invalidates the capture.

    ``loss_fn(model, *inputs) -> (loss, metrics)`` is the body of one micro-step (default: the AVMAE pre-training
    loss); ``metrics`` is a tuple of 0-dim device tensors returned by ``__call__`` next to the gradient norm."""

    def __init__(self, trainer, *examples: torch.Tensor, warmup: int = 3, capture_error_mode: str = "thread_local",
                 overlap_comm: bool = True, loss_fn: Optional[Callable] = None):
        self.trainer = trainer
        self.loss_fn = loss_fn or _avmae_loss
        self.accum_iter = int(trainer.accum_iter)
        self.distributed = trainer.distributed
        self.overlap_comm = self.distributed and overlap_comm and os.environ.get("DAVF_GRAPH_NCCL", "1") != "0"
        self.inputs = [torch.empty_like(e) for e in examples]
        for dst, src in zip(self.inputs, examples):
            dst.copy_(src)
        opt = trainer.optimizer
        saved = self._snapshot()
        # eager warm-up on a side stream: instantiates kernels / attributes / TMA descriptors / allocator pools / NCCL
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(max(1, warmup)):
                for micro in range(self.accum_iter):
                    self._one_step(final=micro == self.accum_iter - 1)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        opt._sync_hp()
        from .. import kernels as K
        self.graph_micro = None
        self.launches_micro = 0
        if self.accum_iter > 1:
            trainer.accums = 0
            self.graph_micro = torch.cuda.CUDAGraph()
            n0 = K.launch_count()
            with torch.cuda.graph(self.graph_micro, capture_error_mode=capture_error_mode):
                self._micro_out = self._one_step(final=False, sync_hp=False, capturing=True)
            self.launches_micro = K.launch_count() - n0
        trainer.accums = self.accum_iter - 1
        self.graph = torch.cuda.CUDAGraph()
        n0 = K.launch_count()
        with torch.cuda.graph(self.graph, capture_error_mode=capture_error_mode):
            self._final_out = self._one_step(final=True, sync_hp=False, capturing=True)
        self.launches_final = K.launch_count() - n0 + (1 if (self.distributed and not self.overlap_comm) else 0)
        # kernels of ONE optimizer step (accum_iter micro-steps)
        self.launches_per_step = self.launches_final + (self.accum_iter - 1) * self.launches_micro
        self.metrics, self.grad_norm = self._final_out
        torch.cuda.synchronize()
        self._restore(saved)
        self._micro = 0

    # -- construction must not train: state snapshot / restore -------------------------------------------
    def _snapshot(self):
        tr = self.trainer
        opt, st = tr.optimizer, tr.store
        bufs = [b for b in tr.model_without_ddp.buffers()]
        return dict(p=st.flat_p.clone(), lp=st.flat_lp.clone(), g=st.flat_g.clone(), m=opt.flat_m.clone(), v=opt.flat_v.clone(),
                    scal=opt.scal.clone(), sumsq=opt.grad_sumsq.clone(), opt_steps=opt.n_steps, tr_steps=tr.n_steps.clone(),
                    accums=tr.accums, bufs=[(b, b.clone()) for b in bufs])

    def _restore(self, s):
        tr = self.trainer
        opt, st = tr.optimizer, tr.store
        with torch.no_grad():
            st.flat_p.copy_(s["p"]); st.flat_lp.copy_(s["lp"]); st.flat_g.copy_(s["g"])
            opt.flat_m.copy_(s["m"]); opt.flat_v.copy_(s["v"]); opt.scal.copy_(s["scal"]); opt.grad_sumsq.copy_(s["sumsq"])
            for b, c in s["bufs"]:
                b.copy_(c)
        st.mark_lowp_fresh()                        # shadows restored together with the parameters
        opt.n_steps = s["opt_steps"]                # (the capture pass ran the host side of end_step(): undo its count)
        opt._open = False
        tr.n_steps.copy_(s["tr_steps"])
        tr.accums = s["accums"]
        if tr.sync is not None:
            tr.sync.reset()
            tr.sync.fuse_optimizer = False
        torch.cuda.synchronize()

    # -- one micro-step (eager warm-up and capture share this body) ----------------------------------------
    def _one_step(self, final: bool, sync_hp: bool = True, capturing: bool = False):
        tr = self.trainer
        loss, metrics = self.loss_fn(tr.model, *self.inputs)
        if not final:                                    # accumulate only (misc.py:144-148: no_sync, no optimizer)
            tr.backward(loss)
            return tuple(metrics), None
        if self.distributed and not self.overlap_comm:
            tr.sync.enabled = False                      # no NCCL inside the captured region
            loss.backward()
            tr.store.join_side_streams(torch.cuda.current_stream())
            tr.accums += 1
            if not capturing:
                self._reduce_and_step(sync_hp)
            return tuple(metrics), tr.optimizer.grad_norm()
        n0 = int(tr.n_steps)
        tr._step(loss, sync_hp=sync_hp)                  # backward (+ bucket all-reduce) with the optimizer applied bucket by bucket
        tr.n_steps.fill_(n0)                              # __call__ counts the steps
        return tuple(metrics), tr.optimizer.grad_norm()

    def _reduce_and_step(self, sync_hp: bool = True):
        import torch.distributed as dist
        tr = self.trainer
        dist.all_reduce(tr.store.flat_g, op=dist.ReduceOp.SUM)      # 1/world is folded into AdamW's grad scale
        tr.optimizer.step(zero_grad=True, sync_hp=sync_hp)
        tr.accums = 0

    # -- inputs --------------------------------------------------------------------------------------------
    def _load_inputs(self, inputs: Sequence[torch.Tensor]) -> None:
        """Inputs may live on the host (pinned).  Host inputs travel on a dedicated copy stream into one of two
        staging sets and reach the graph's static inputs with device-to-device copies, so the 45 MB H2D transfer of
        step t+1 (~0.9 ms over PCIe) overlaps the compute of step t whenever the host runs ahead of the GPU."""
        assert len(inputs) == len(self.inputs)
        if all(t.device.type == "cpu" for t in inputs):
            cur = torch.cuda.current_stream()
            if not hasattr(self, "_cstream"):
                self._cstream = torch.cuda.Stream()
                self._stage = [[torch.empty_like(t) for t in self.inputs] for _ in range(2)]
                self._stage_free = [None, None]
                self._k = 0
            k = self._k
            self._k ^= 1
            if self._stage_free[k] is not None:
                self._cstream.wait_event(self._stage_free[k])       # the D2D copies that last read this staging slot
            with torch.cuda.stream(self._cstream):
                for dst, src in zip(self._stage[k], inputs):
                    dst.copy_(src, non_blocking=True)
                ready = self._cstream.record_event()
            cur.wait_event(ready)
            for dst, src in zip(self.inputs, self._stage[k]):
                dst.copy_(src, non_blocking=True)
            self._stage_free[k] = cur.record_event()
        else:
            for dst, src in zip(self.inputs, inputs):
                dst.copy_(src, non_blocking=True)

    def __call__(self, *inputs: torch.Tensor):
        """One MICRO-step.  Returns (*metrics, grad_norm); grad_norm is None on accumulate-only micro-steps (the
        first accum_iter - 1 of every optimizer step).  For the default loss: (loss_image, loss_audio, grad_norm)."""
        self._load_inputs(inputs)
        tr = self.trainer
        if self._micro < self.accum_iter - 1:
            self.graph_micro.replay()
            self._micro += 1
            tr.accums += 1
            return (*self._micro_out[0], None)
        tr.optimizer._sync_hp()
        self.graph.replay()
        self._micro = 0
        tr.accums = 0
        if self.distributed and not self.overlap_comm:
            self._reduce_and_step(sync_hp=False)
            self.grad_norm = tr.optimizer.grad_norm()
        else:
            tr.optimizer.n_steps += 1
        tr.n_steps += 1
        return (*self.metrics, self.grad_norm)

    # -- pipelined host read-back -------------------------------------------------------------------------
    def step_async(self, *inputs: torch.Tensor) -> None:
        """One micro-step whose (*metrics, grad_norm) are copied to pinned host memory asynchronously, to be
        collected with ``pop_metrics``.  A training loop that logs step t's loss after it has launched step t+1
        (train.py:166 reads ``loss.item()`` for the meters only) never drains the GPU: the host-side graph launch
        of the next step (~0.7 ms for ~1,100 kernel nodes) overlaps the current step instead of following it."""
        out = self(*inputs)
        n = len(out)
        if not hasattr(self, "_ring"):
            self._ring = [(torch.empty(n, dtype=torch.float32).pin_memory(), torch.cuda.Event()) for _ in range(4)]
            self._devn = torch.empty(n, dtype=torch.float32, device=self.inputs[0].device)
            self._nan = torch.full((), float("nan"), device=self.inputs[0].device)
            self._head = self._tail = 0
        assert self._head - self._tail < len(self._ring), "pop_metrics() must be called to drain the ring"
        torch.stack([(self._nan if t is None else t.reshape(()).float()) for t in out], out=self._devn)
        host, ev = self._ring[self._head % len(self._ring)]
        host.copy_(self._devn, non_blocking=True)
        ev.record()
        self._head += 1

    def pending(self) -> int:
        return getattr(self, "_head", 0) - getattr(self, "_tail", 0)

    def pop_metrics(self):
        """(*metrics, grad_norm) of the oldest micro-step not yet collected, as Python floats (grad_norm is nan on
        accumulate-only micro-steps); waits for that step only."""
        assert self.pending() > 0
        host, ev = self._ring[self._tail % len(self._ring)]
        ev.synchronize()
        self._tail += 1
        return tuple(float(x) for x in host)

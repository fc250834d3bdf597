"""Whole-step CUDA graph: forward + backward (+ bucket all-reduce) + fused AdamW captured once and
replayed, removing the ~2,000 per-step launches' CPU cost and every host sync from the loop
(eager PyTorch on this path is launch-bound).  Hyper-parameters (lr / weight decay,
beta^t, grad scale) live in device tables the captured kernels read, so LR schedules need no
re-capture; the mask noise comes from torch's graph-safe Philox generator state.

Data-parallel runs capture the bucketed gradient all-reduce too (``overlap_comm``, the default): every bucket's
NCCL call is recorded on the communication stream at the point of backward where its last gradient has been
produced, so in the replayed graph the all-reduce of the decoder / late-encoder buckets runs under the rest of
backward, and the fused AdamW node follows the last bucket.  NCCL needs its communicator and internal stream
to exist before capture: the eager warm-up steps run the identical bucketed path.  ``DAVF_GRAPH_NCCL=0`` (or
``overlap_comm=False``) falls back to capturing forward + backward only, with one all-reduce over the flat
gradient buffer and the AdamW launch issued after the replay."""
from __future__ import annotations

import os

import torch


class GraphedTrainStep:
    """Capture once, replay per step.  Drop references to the outputs (losses, predictions) of earlier EAGER
    steps before constructing this object: they carry ``record_stream`` marks from the multi-stream forward, and
    freeing such a tensor while a capture is in progress invalidates the capture."""

    def __init__(self, trainer, image_example: torch.Tensor, audio_example: torch.Tensor, warmup: int = 3,
                 capture_error_mode: str = "thread_local", overlap_comm: bool = True):
        assert trainer.accum_iter == 1, "graph capture covers one full optimizer step (accum_iter == 1)"
        self.trainer = trainer
        self.distributed = trainer.sync is not None
        self.overlap_comm = self.distributed and overlap_comm and os.environ.get("DAVF_GRAPH_NCCL", "1") != "0"
        self.image = torch.empty_like(image_example)
        self.audio = torch.empty_like(audio_example)
        self.image.copy_(image_example)
        self.audio.copy_(audio_example)
        opt = trainer.optimizer
        # eager warm-up on a side stream: instantiates kernels / attributes / TMA descriptors / allocator pools
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                self._one_step()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        opt._sync_hp()
        self.graph = torch.cuda.CUDAGraph()
        from .. import kernels as K
        n0 = K.launch_count()
        with torch.cuda.graph(self.graph, capture_error_mode=capture_error_mode):
            self.loss_image, self.loss_audio, self.grad_norm = self._one_step(sync_hp=False, capturing=True)
        self.launches_per_step = K.launch_count() - n0 + (1 if (self.distributed and not self.overlap_comm) else 0)

    def _one_step(self, sync_hp: bool = True, capturing: bool = False):
        tr = self.trainer
        li, la, _, _ = tr.model(self.image, self.audio)
        if self.distributed and not self.overlap_comm:
            tr.sync.enabled = False                      # no NCCL inside the captured region
            (li + la).backward()
            tr.store.join_side_streams(torch.cuda.current_stream())
            if not capturing:
                self._reduce_and_step(sync_hp)
            return li.detach(), la.detach(), tr.optimizer.grad_norm()
        n0 = int(tr.n_steps)
        tr._step(li + la, sync_hp=sync_hp)               # backward (+ bucket all-reduce) with the optimizer applied bucket by bucket
        tr.n_steps.fill_(n0)                              # __call__ counts the steps
        return li.detach(), la.detach(), tr.optimizer.grad_norm()

    def _reduce_and_step(self, sync_hp: bool = True):
        import torch.distributed as dist
        tr = self.trainer
        dist.all_reduce(tr.store.flat_g, op=dist.ReduceOp.SUM)      # 1/world is folded into AdamW's grad scale
        tr.optimizer.step(zero_grad=True, sync_hp=sync_hp)
        tr.accums = 0

    def __call__(self, image: torch.Tensor, audio: torch.Tensor):
        """image / audio may live on the host (pinned).  Host inputs travel on a dedicated copy stream into one of two
        staging buffers and reach the graph's static input with a device-to-device copy, so the 45 MB H2D transfer of
        step t+1 (~0.9 ms over PCIe) overlaps the compute of step t whenever the host runs ahead of the GPU."""
        if image.device.type == "cpu" and audio.device.type == "cpu":
            cur = torch.cuda.current_stream()
            if not hasattr(self, "_cstream"):
                self._cstream = torch.cuda.Stream()
                self._stage = [(torch.empty_like(self.image), torch.empty_like(self.audio)) for _ in range(2)]
                self._stage_free = [None, None]
                self._k = 0
            k = self._k
            self._k ^= 1
            si, sa = self._stage[k]
            if self._stage_free[k] is not None:
                self._cstream.wait_event(self._stage_free[k])       # the D2D copy that last read this staging slot
            with torch.cuda.stream(self._cstream):
                si.copy_(image, non_blocking=True)
                sa.copy_(audio, non_blocking=True)
                ready = self._cstream.record_event()
            cur.wait_event(ready)
            self.image.copy_(si, non_blocking=True)
            self.audio.copy_(sa, non_blocking=True)
            self._stage_free[k] = cur.record_event()
        else:
            self.image.copy_(image, non_blocking=True)
            self.audio.copy_(audio, non_blocking=True)
        self.trainer.optimizer._sync_hp()
        self.graph.replay()
        if self.distributed and not self.overlap_comm:
            self._reduce_and_step(sync_hp=False)
            self.grad_norm = self.trainer.optimizer.grad_norm()
        else:
            self.trainer.optimizer.n_steps += 1
        self.trainer.n_steps += 1
        return self.loss_image, self.loss_audio, self.grad_norm

    # -- pipelined host read-back -------------------------------------------------------------------------
    def step_async(self, image: torch.Tensor, audio: torch.Tensor) -> None:
        """One step whose (loss_image, loss_audio, grad_norm) are copied to pinned host memory asynchronously, to be
        collected with ``pop_metrics``.  A training loop that logs step t's loss after it has launched step t+1
        (train.py:166 reads ``loss.item()`` for the meters only) never drains the GPU: the host-side graph launch
        of the next step (~0.7 ms for ~1,100 kernel nodes) overlaps the current step instead of following it."""
        if not hasattr(self, "_ring"):
            self._ring = [(torch.empty(3, dtype=torch.float32).pin_memory(), torch.cuda.Event()) for _ in range(4)]
            self._dev3 = torch.empty(3, dtype=torch.float32, device=self.image.device)
            self._head = self._tail = 0
        assert self._head - self._tail < len(self._ring), "pop_metrics() must be called to drain the ring"
        li, la, norm = self(image, audio)
        torch.stack((li.reshape(()), la.reshape(()), norm.reshape(())), out=self._dev3)
        host, ev = self._ring[self._head % len(self._ring)]
        host.copy_(self._dev3, non_blocking=True)
        ev.record()
        self._head += 1

    def pending(self) -> int:
        return getattr(self, "_head", 0) - getattr(self, "_tail", 0)

    def pop_metrics(self):
        """(loss_image, loss_audio, grad_norm) of the oldest step not yet collected, as Python floats; waits for that
        step only."""
        assert self.pending() > 0
        host, ev = self._ring[self._tail % len(self._ring)]
        ev.synchronize()
        self._tail += 1
        return float(host[0]), float(host[1]), float(host[2])

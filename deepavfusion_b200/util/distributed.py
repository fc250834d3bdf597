"""Data-parallel plumbing (mirror of the hot-path parts of reference util/distributed.py:66-100 and
the DDP wrap at util/misc.py:32-34): one process per GPU, ``torch.distributed`` over NCCL / NVLink,
and a bucketed gradient all-reduce that overlaps backward.

The path shards over the batch; the only exchange step is the sum-all-reduce of the
flat f32 gradient buffer once per optimizer step.  Buckets are contiguous ranges of that buffer cut
from its END (parameters are laid out in forward order, so backward finishes them first); a bucket is
launched on a side stream as soon as the last backward region touching it has enqueued its kernels.
Division by world size is folded into the fused AdamW's grad_scale.
"""
from __future__ import annotations

import os
from typing import List, Optional

import torch
import torch.distributed as dist

from ..params import ALIGN, ParamStore


def is_dist_avail_and_initialized() -> bool:
    return dist.is_available() and dist.is_initialized()


def get_world_size() -> int:
    return dist.get_world_size() if is_dist_avail_and_initialized() else 1


def get_rank() -> int:
    return dist.get_rank() if is_dist_avail_and_initialized() else 0


def init_from_env(backend: Optional[str] = None) -> int:
    """torchrun-style init (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*).  The reference launcher's
    NCCL_P2P_DISABLE=1 / NCCL_P2P_LEVEL=LOC (launcher.py:74-76) must NOT be inherited on an NVSwitch box."""
    for bad in ("NCCL_P2P_DISABLE", "NCCL_P2P_LEVEL"):
        os.environ.pop(bad, None)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        return 0
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if backend == "nccl":
        torch.cuda.set_device(local)
    if not dist.is_initialized():
        kw = {}
        if backend == "nccl":
            kw["device_id"] = torch.device("cuda", local)
            if os.environ.get("DAVF_NCCL_HIGH_PRIORITY", "0") == "1":
                # opt-in (see nccl_registered_zeros): the all-reduce CTAs go ahead of queued compute CTAs
                kw["pg_options"] = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
        dist.init_process_group(backend=backend, **kw)
    return local


def nccl_registered_zeros(numel: int, device, group=None):
    """An f32 buffer from NCCL's own allocator (``ncclMemAlloc``), registered with the communicator, so that NVLS
    all-reduces work on it in place (multimem load-reduce / store straight on the user buffer) instead of staging
    every bucket through NCCL's internal buffers.  Returns ``(tensor, pool)`` -- keep the pool alive as long as the
    tensor -- or ``(None, reason)`` when the process group is not NCCL or registration is unavailable.

    OPT-IN (``DAVF_NCCL_REGISTER=1``, usually with ``DAVF_NCCL_HIGH_PRIORITY=1``).  Measured on 8 B200s: the 64 MB bucket
    all-reduce drops from 24 NVLS channels / 0.38 ms to 8 channels / 0.28 ms and the step from 18.19 to 16.74 ms
    (profiles/r2/bench_n8_registered.json vs bench_n8_x.json).  It is not the default because one of the four 8-GPU jobs
    run with it (the only one with DAVF_COMM_SMS=8) hung without output and the GPU budget ended before the cause could
    be isolated; the unregistered path has never hung in any run."""
    if os.environ.get("DAVF_NCCL_REGISTER", "0") != "1":
        return None, "off (opt in with DAVF_NCCL_REGISTER=1)"
    if not (is_dist_avail_and_initialized() and dist.get_backend(group) == "nccl"):
        return None, "process group is not NCCL"
    try:
        pg = group if group is not None else dist.distributed_c10d._get_default_group()
        backend = pg._get_backend(torch.device(device))
        pool = torch.cuda.MemPool(backend.mem_allocator)
        with torch.cuda.use_mem_pool(pool, device=torch.device(device)):
            buf = torch.zeros(numel, dtype=torch.float32, device=device)
        backend.register_mem_pool(pool)
        return buf, pool
    except Exception as e:      # noqa: BLE001 -- an unregistered buffer is slower, not wrong; the reason is reported by bench.py
        return None, f"{type(e).__name__}: {e}"


class GradSync:
    """Bucketed, backward-overlapped sum-all-reduce of ``store.flat_g`` (replaces DDP's reducer).

    With ``fuse_optimizer`` set for a backward pass (``Trainer.step`` does it: the optimizer step follows
    immediately), each bucket's fused AdamW range launch follows its all-reduce -- on a stream of its own, so that
    bucket b's AdamW runs under bucket b+1's all-reduce instead of in series with it (measured at N = 2: the
    all-reduce -> AdamW chain, 0.21 + 0.13 ms per bucket on one stream, was the backlogged critical path from the
    middle of backward to the end of the step) -- and the HBM-bound optimizer (8.97 GB of traffic per step) runs
    under the tensor-core-bound rest of backward instead of after it.  Works with world size 1 too (no all-reduce, only the overlapped optimizer)."""

    def __init__(self, store: ParamStore, bucket_mb: float = 64.0, group=None, optimizer=None, tail_mb: Optional[float] = None):
        self.store = store
        self.group = group
        self.world = dist.get_world_size(group) if is_dist_avail_and_initialized() else 1
        self.optimizer = optimizer                # FusedAdamW or None
        self.fuse_optimizer = False               # set per backward pass by Trainer.step
        self.enabled = True                       # False while accumulating (DDP no_sync, misc.py:144-148)
        target = int(bucket_mb * 1024 * 1024 / 4)
        tail_mb = float(os.environ.get("DAVF_TAIL_BUCKET_MB", bucket_mb if tail_mb is None else tail_mb))
        tail = max(1, min(target, int(tail_mb * 1024 * 1024 / 4)))
        n = len(store.params)
        # buckets of parameter indices, cut from the end of the buffer (the order backward produces them in).  ``tail_mb``
        # (default: same as bucket_mb) cuts the last two buckets' worth -- the first layers, whose gradients arrive when
        # nothing is left to hide an exchange under -- finer; measured at N = 8 it did not pay (17.05 vs 16.74 ms: the
        # extra NVLS launches cost more than the shorter tail saves), so it is off.
        self.buckets: List[List[int]] = []
        cur: List[int] = []
        size = 0
        for k in range(n - 1, -1, -1):
            b, e = store.span(k)
            cur.append(k)
            size += e - b
            if size >= (tail if b < 2 * target else target) and b % ALIGN == 0:      # (a stacked k / v pair may start off the grid: never cut there)
                self.buckets.append(cur)
                cur, size = [], 0
        if cur:
            self.buckets.append(cur)
        self.bucket_of = {}
        self.ranges = []
        for bi, ks in enumerate(self.buckets):
            lo = min(store.span(k)[0] for k in ks)
            hi = max(store.span(k)[1] for k in ks)
            self.ranges.append((lo, hi))
            for k in ks:
                self.bucket_of[k] = bi
        self.trainable = [sum(1 for k in ks if store.params[k].requires_grad) for ks in self.buckets]
        self.cuda = store.flat_g.is_cuda
        self.comm_stream = torch.cuda.Stream() if self.cuda else None
        self.opt_stream = torch.cuda.Stream() if self.cuda else None
        self.reset()
        store.sync = self

    def reset(self):
        self.pending = list(self.trainable)
        self.seen = set()
        self.launched = [False] * len(self.buckets)
        self.handles = []

    def _active(self) -> bool:
        return self.enabled and (self.world > 1 or self.fuse_optimizer)

    def params_done(self, idxs):
        if not self._active():
            return
        for k in idxs:
            if k in self.seen:
                continue
            self.seen.add(k)
            bi = self.bucket_of[k]
            self.pending[bi] -= 1
            if self.pending[bi] == 0:
                self._launch(bi)

    def _launch(self, bi: int):
        if self.launched[bi]:
            return
        self.launched[bi] = True
        lo, hi = self.ranges[bi]
        buf = self.store.flat_g[lo:hi]
        if self.cuda:
            # the bucket's gradients were produced on several streams (image / audio / fusion branches run on
            # their own streams): wait for everything enqueued so far on each of them
            waits = [torch.cuda.current_stream()] + list(self.store.side_streams)
            if self.store.main_stream is not None:
                waits.append(self.store.main_stream)
            for s in waits:
                self.comm_stream.wait_stream(s)
            with torch.cuda.stream(self.comm_stream):
                if self.world > 1:
                    dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group)
            if self.fuse_optimizer:
                self.opt_stream.wait_stream(self.comm_stream)      # this bucket's all-reduce (and everything it waited for)
                with torch.cuda.stream(self.opt_stream):
                    self.optimizer.step_range(lo, hi)
        else:
            if self.world > 1:
                self.handles.append((dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group, async_op=True), lo, hi))
            elif self.fuse_optimizer:
                self.optimizer.step_range(lo, hi)

    def finish(self):
        """End of backward: launch whatever is left (gradients that arrive through autograd itself, e.g.
        ``fusion_tokens``), then make the compute stream wait for the communication stream."""
        if not self._active():
            return
        for bi in range(len(self.buckets)):
            self._launch(bi)
        if self.cuda:
            torch.cuda.current_stream().wait_stream(self.comm_stream)
            torch.cuda.current_stream().wait_stream(self.opt_stream)
        for h, lo, hi in self.handles:
            h.wait()
            if self.fuse_optimizer:
                self.optimizer.step_range(lo, hi)
        if self.fuse_optimizer:
            self.optimizer.end_step()
            self.fuse_optimizer = False
        self.reset()

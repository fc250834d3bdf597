"""Flat parameter / gradient / bf16-shadow storage laid out for 180 GB of HBM.

All parameters of the top-level module become views into ONE f32 buffer (so a single fused
AdamW launch and a handful of large gradient buckets cover the model), gradients are views into
one f32 buffer that the wgrad GEMMs accumulate into directly (no autograd ``AccumulateGrad``
copies; the ``+=`` also implements gradient accumulation over micro-steps, misc.py:144-148), and
every parameter has a bf16 shadow at the same offset that the tensor-core kernels read (what
autocast re-casts on every forward in the reference).

Parameters stay ordinary ``nn.Parameter`` objects: ``state_dict()``, ``named_parameters()``,
``load_state_dict(strict=True)``, ``torch.optim`` and ``param.grad`` consumers (misc.py:151-163)
keep working.  Order inside the buffers is the order of first use in the forward pass, so that
gradient buckets (contiguous ranges) complete in reverse order during backward.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch
from torch import nn

ALIGN = 64          # elements; keeps every tensor 256-byte (f32) / 128-byte (bf16) aligned for TMA


import re

_BLK = re.compile(r"^(?:encoder\.)?(image|audio)\.blocks\.(\d+)\.")
_FUS = re.compile(r"^(?:encoder\.)?fusion_blocks\.(\d+)\.")
_DEC = re.compile(r"^(image|audio)_decoder_(\w+?)(?:\.(\d+))?(?:\.|$)")


def forward_order_key(name: str):
    """Sort key = position of the parameter's FIRST use in the forward pass (deepavfusion.py:88-113,
    avmae.py:216-236).  Backward produces gradients in the reverse order, so gradient buckets cut
    from the END of the flat buffer complete early and their all-reduce overlaps the rest of
    backward (what DDP's bucket order does in the reference, misc.py:34)."""
    base = name[len("encoder."):] if name.startswith("encoder.") else name
    if base == "fusion_tokens":
        return (0, 0, 0)
    m = _BLK.match(name)
    if m:
        return (2, int(m.group(2)), 0 if m.group(1) == "image" else 1)
    m = _FUS.match(name)
    if m:
        return (2, int(m.group(1)), 2)
    if base.startswith(("image.patch_embed", "image.pos_embed", "image.cls_token")):
        return (1, 0, 0)
    if base.startswith(("audio.patch_embed", "audio.pos_embed", "audio.cls_token")):
        return (1, 1, 0)
    if base.startswith(("image.norm", "audio.norm", "fusion_norm")):
        return (3, 0, 0)
    m = _DEC.match(name)
    if m:
        stage = 4 if m.group(1) == "image" else 5
        what = m.group(2)
        if what == "blocks":
            return (stage, 1 + int(m.group(3)), 0)
        return (stage, 0 if what in ("embed", "mask_token", "pos_embed") else 100, 0)
    return (9, 0, 0)


_PAIR_KV = {"attn.k.weight": 1, "attn.v.weight": 2, "attn.k.bias": 3, "attn.v.bias": 4}


def _pair_kv_rank(name: str) -> int:
    """Inside a fusion block the pair-attention ``k`` and ``v`` Linears (fusion_blocks.py:228-229) read the
    same input, so their weights (and biases) are stored back to back: [k.weight; v.weight] is then ONE
    [qk + D, 2D] matrix and k / v become a single GEMM in forward, dgrad and wgrad (``ParamStore.stacked``)."""
    if _FUS.match(name):
        for suffix, r in _PAIR_KV.items():
            if name.endswith(suffix):
                return r
    return 0


class ParamStore:
    def __init__(self, module: nn.Module):
        named = [(n, p) for n, p in module.named_parameters()]
        assert named, "module has no parameters"
        dev = named[0][1].device
        order = sorted(range(len(named)), key=lambda i: (forward_order_key(named[i][0]), _pair_kv_rank(named[i][0]), i))
        self.names: List[str] = []
        self.params: List[nn.Parameter] = []
        self.offsets: List[int] = []
        off = 0
        end_prev, rank_prev = 0, 0
        for i in order:
            n, p = named[i]
            assert p.dtype == torch.float32, f"{n}: master parameters must be f32"
            rank = _pair_kv_rank(n)
            if rank in (2, 4) and rank_prev == rank - 1 and end_prev % 8 == 0:
                off = end_prev           # v directly behind k (no alignment gap): [k; v] is one stacked tensor
            self.names.append(n); self.params.append(p); self.offsets.append(off)
            end_prev, rank_prev = off + p.numel(), rank
            off = (end_prev + ALIGN - 1) // ALIGN * ALIGN
        self.numel = off
        self.device = dev
        self.flat_p = torch.zeros(off, dtype=torch.float32, device=dev)
        self.flat_g = torch.zeros(off, dtype=torch.float32, device=dev)
        self.flat_lp = torch.zeros(off, dtype=torch.bfloat16, device=dev)
        self._index: Dict[int, int] = {}
        self._lp: List[torch.Tensor] = []
        self._g: List[torch.Tensor] = []
        with torch.no_grad():
            for k, (p, o) in enumerate(zip(self.params, self.offsets)):
                view = self.flat_p[o:o + p.numel()].view(p.shape)
                view.copy_(p.data)
                p.data = view
                self._index[id(p)] = k
                self._lp.append(self.flat_lp[o:o + p.numel()].view(p.shape))
                self._g.append(self.flat_g[o:o + p.numel()].view(p.shape))
                if p.requires_grad:
                    p.grad = self._g[k]
        self._versions = None
        self._depth = 0
        self.sync = None          # optional util.distributed.GradSync (data-parallel bucket all-reduce)
        self.side_streams = []    # streams the model issues independent branches on (see DeepAVFusion._side_streams)
        self.main_stream = None   # the stream the current forward was issued on
        self._join_pending = False
        self._lowp_grads = {}     # data_ptr of an f32 activation gradient -> (tensor, version, bf16 copy); see stash_lowp_grad
        self.refresh_lowp(force=True)

    # -- re-entrancy: nested module forwards skip the per-forward checks ------------------------
    def __enter__(self):
        if self._depth == 0:
            self._lowp_grads.clear()
        self._depth += 1
        return self

    def __exit__(self, *exc):
        self._depth -= 1
        return False

    # -- identity ---------------------------------------------------------------------------
    def owns(self, module: nn.Module) -> bool:
        """True while every parameter still aliases the flat buffer (``.to()`` / ``.cuda()`` /
        re-assignment of ``.data`` break the aliasing and require a rebuild)."""
        for p, o in zip(self.params, self.offsets):
            if p.data_ptr() != self.flat_p.data_ptr() + 4 * o:
                return False
        return True

    def covers(self, module: nn.Module) -> bool:
        """Cheap per-forward check: a sample of ``module``'s parameters is in this store and still
        aliases the flat buffer."""
        ps = list(module.parameters())
        for p in (ps[0], ps[len(ps) // 2], ps[-1]):
            k = self._index.get(id(p))
            if k is None or p.data_ptr() != self.flat_p.data_ptr() + 4 * self.offsets[k]:
                return False
        return True

    # -- bf16 shadows -------------------------------------------------------------------------
    def _version_key(self):
        return tuple(p._version for p in self.params)

    def refresh_lowp(self, force: bool = False) -> bool:
        """Re-cast f32 -> bf16 if any parameter was modified in place since the last cast
        (optimizer step, load_state_dict, init).  One launch over the flat buffer."""
        key = self._version_key()
        if not force and key == self._versions:
            return False
        from . import kernels as K
        K.cast_flat_bf16(self.flat_p, self.flat_lp)
        self._versions = key
        return True

    def mark_lowp_fresh(self) -> None:
        """Called by the fused optimizer, which writes the bf16 shadows itself."""
        self._versions = self._version_key()

    def lowp(self, p: nn.Parameter) -> torch.Tensor:
        return self._lp[self._index[id(p)]]

    def stacked(self, p1: nn.Parameter, p2: nn.Parameter):
        """(bf16 shadow, gradient) views of ``[p1; p2]`` as one tensor stacked along dim 0.  Requires the two
        parameters to be adjacent in the flat buffers (see ``_pair_kv_rank``) with equal trailing shapes."""
        k1, k2 = self._index[id(p1)], self._index[id(p2)]
        o1, o2 = self.offsets[k1], self.offsets[k2]
        if o2 != o1 + p1.numel() or p1.shape[1:] != p2.shape[1:]:
            raise RuntimeError(f"parameters {self.names[k1]} / {self.names[k2]} are not stacked in the flat buffer "
                               f"(offsets {o1}+{p1.numel()} vs {o2}); build the ParamStore over the whole model")
        shape = (p1.shape[0] + p2.shape[0],) + tuple(p1.shape[1:])
        n = p1.numel() + p2.numel()
        if (p1.requires_grad and not self._attached(p1, k1)) or (p2.requires_grad and not self._attached(p2, k2)):
            self.ensure_grads(force=True)
        return self.flat_lp[o1:o1 + n].view(shape), self.flat_g[o1:o1 + n].view(shape)

    # -- bf16 copies of activation gradients ------------------------------------------------------
    def stash_lowp_grad(self, g: torch.Tensor, lp: torch.Tensor) -> None:
        """The kernel that produced the f32 gradient ``g`` (LayerNorm backward) also wrote its bf16 copy ``lp``; the
        next backward region, which needs ``g`` as a bf16 GEMM operand, picks it up with ``take_lowp_grad`` instead
        of launching a cast.  The entry pins ``g``, so a later tensor with the same address is the same tensor."""
        self._lowp_grads[g.data_ptr()] = (g, g._version, lp)

    def take_lowp_grad(self, g: torch.Tensor):
        e = self._lowp_grads.pop(g.data_ptr(), None)
        if e is None or e[0].numel() != g.numel() or e[0]._version != e[1] or g.dtype != torch.float32:
            return None
        return e[2]

    # -- gradients ------------------------------------------------------------------------------
    def _attached(self, p: nn.Parameter, k: int) -> bool:
        g = p.grad
        return g is not None and g.data_ptr() == self._g[k].data_ptr() and g.dtype == torch.float32

    def ensure_grads(self, force: bool = False) -> None:
        """Re-attach ``.grad`` views after a ``zero_grad(set_to_none=True)`` by a stock optimizer
        (torch's default): one flat memset instead of ~900 small ones."""
        first = next((p for p in self.params if p.requires_grad), None)
        if first is None or (not force and self._attached(first, self._index[id(first)])):
            return
        if all(p.grad is None for p in self.params if p.requires_grad):
            self.flat_g.zero_()
        for k, p in enumerate(self.params):
            if p.requires_grad:
                if p.grad is not None and not self._attached(p, k):
                    self._g[k].copy_(p.grad)
                p.grad = self._g[k]

    def rehome_grads(self, buf: torch.Tensor) -> None:
        """Moves ``flat_g`` into ``buf`` (same size / dtype / device; e.g. memory registered with the NCCL communicator,
        util.distributed.nccl_registered_zeros) and re-points every gradient view at it."""
        assert buf.shape == self.flat_g.shape and buf.dtype == self.flat_g.dtype and buf.device == self.flat_g.device
        with torch.no_grad():
            buf.copy_(self.flat_g)
            self.flat_g = buf
            self._g = [buf[o:o + p.numel()].view(p.shape) for p, o in zip(self.params, self.offsets)]
            for k, p in enumerate(self.params):
                if p.requires_grad:
                    p.grad = self._g[k]
        self._lowp_grads.clear()

    def grad(self, p: nn.Parameter) -> torch.Tensor:
        k = self._index[id(p)]
        if not self._attached(p, k):
            self.ensure_grads(force=True)
        return self._g[k]

    def zero_grad(self) -> None:
        self.flat_g.zero_()
        for k, p in enumerate(self.params):
            if p.requires_grad and not self._attached(p, k):
                p.grad = self._g[k]

    def index_of(self, p: nn.Parameter) -> int:
        return self._index[id(p)]

    def span(self, k: int) -> Tuple[int, int]:
        """[begin, end) element range of parameter k in the flat buffers (end includes alignment padding)."""
        end = self.offsets[k + 1] if k + 1 < len(self.offsets) else self.numel
        return self.offsets[k], end

    def done(self, params) -> None:
        """Called at the end of a backward region: the kernels producing these parameters' gradients
        are all enqueued.  Drives the overlapped bucket all-reduce when data-parallel."""
        if self.sync is not None:
            self.sync.params_done([self._index[id(p)] for p in params if p.requires_grad])
        self._schedule_join()

    def _schedule_join(self) -> None:
        """Parameter gradients are written by our kernels directly into ``flat_g`` on whatever stream the
        backward node runs on -- autograd does not know about them, so it does not order them before the
        caller's stream at the end of ``backward()``.  Queue (once per backward pass) an engine callback that
        makes the caller's stream wait for the side streams before anything (optimizer, all-reduce, grad
        norm) reads the gradients."""
        if not self.side_streams or self._join_pending:
            return
        self._join_pending = True

        def join():
            self._join_pending = False
            self._lowp_grads.clear()
            self.join_side_streams(self.main_stream)
        torch.autograd.Variable._execution_engine.queue_callback(join)

    def wgrad_stream(self):
        """Dedicated stream for weight-gradient GEMMs (None on CPU tensors or with DAVF_WGRAD_STREAM=0)."""
        ws = self.__dict__.get("_wgrad_stream", False)
        if ws is False:
            import os
            ws = None
            if self.flat_g.is_cuda and os.environ.get("DAVF_WGRAD_STREAM", "1") != "0" and os.environ.get("DAVF_STREAMS", "1") != "0":
                ws = torch.cuda.Stream(self.flat_g.device)
                self.side_streams.append(ws)
            self.__dict__["_wgrad_stream"] = ws
        return ws

    def aux_streams(self, n: int = 2):
        """``n`` auxiliary streams for independent small launches INSIDE one backward / forward region (the fusion block's
        three LayerNorms, its two cross attentions, the two halves of the pair attention): forked from and joined back
        into the region's own stream, so in the captured step graph they are parallel branches and the region's
        dependent chain -- the critical path of every encoder layer -- gets shorter.  None on CPU tensors or with
        DAVF_STREAMS=0 / DAVF_AUX_STREAMS=0."""
        aux = self.__dict__.get("_aux_streams", False)
        if aux is False:
            import os
            aux = None
            if self.flat_g.is_cuda and os.environ.get("DAVF_STREAMS", "1") != "0" and os.environ.get("DAVF_AUX_STREAMS", "1") != "0":
                aux = [torch.cuda.Stream(self.flat_g.device) for _ in range(n)]
                self.side_streams.extend(aux)
            self.__dict__["_aux_streams"] = aux
        return aux

    def join_side_streams(self, stream=None) -> None:
        """Make ``stream`` (default: the stream the last forward was issued on) wait for the side streams.
        The engine callback may run on an autograd worker thread whose *current* stream is not the caller's,
        so the caller's stream is remembered at forward time rather than queried here."""
        if not self.side_streams:
            return
        cur = stream if stream is not None else (self.main_stream or torch.cuda.current_stream())
        for s in self.side_streams:
            if s != cur:
                cur.wait_stream(s)

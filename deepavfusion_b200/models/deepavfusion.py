"""Early-fusion audio-visual encoder (mirror of reference models/deepavfusion.py:6-118).

Drop-in: same constructor keywords, ``forward(image, audio, image_ids_keep=None,
audio_ids_keep=None, return_embs=False)``, ``embed_dim``, ``image`` / ``audio`` / ``fusion_blocks`` /
``fusion_tokens`` / ``fusion_norm`` attributes, ``params_layer_ids()``, ``load_checkpoint()`` and
``state_dict`` layout.  The arithmetic is the fused sm_100a path of ``functional``.
"""
from __future__ import annotations

from functools import partial

import torch
from torch import nn

import contextlib
import os
from types import SimpleNamespace

from .. import functional as Fn
from . import fusion_blocks, vits
from .layers import FinalNorm, ensure_store
from .vits import _xavier_linear_


class DeepAVFusion(nn.Module):
    def __init__(self,
                 image_arch="vit_base", image_pretrained=True, image_size=(224, 224),
                 audio_arch="vit_base", audio_pretrained=True, audio_size=(128, 192),
                 fusion_arch="factorized_mmi", fusion_layers="all", num_fusion_tkns=(4, 8, 4),
                 fusion_mlp_ratio=1.0, fusion_attn_ratio=0.25, fusion_num_heads=12,
                 drop_path=0.0, attn_drop=0.0, drop=0.0):
        super().__init__()
        self.image = vits.__dict__[image_arch](pretrained=image_pretrained, input_size=image_size, in_chans=3,
                                               use_cls_token=False, drop_path=drop_path, attn_drop=attn_drop, drop=drop)
        self.audio = vits.__dict__[audio_arch](pretrained=audio_pretrained, input_size=audio_size, in_chans=1,
                                               use_cls_token=False, drop_path=drop_path, attn_drop=attn_drop, drop=drop)
        self.embed_dim = self.image.embed_dim
        self.fusion_arch = fusion_arch
        self.num_fusion = tuple(num_fusion_tkns)
        self.fusion_tokens = nn.Parameter(torch.zeros(1, sum(num_fusion_tkns), self.embed_dim))

        if fusion_arch != "factorized_mmi":
            raise NotImplementedError(
                f"fusion_arch={fusion_arch!r}: only 'factorized_mmi' (the arch of every reference config) is built")
        FusionBlock = partial(fusion_blocks.FusionBlock_FactorizedAVInteractions, fusion_tkns=self.num_fusion)
        max_depth = max(len(self.image.blocks), len(self.audio.blocks))
        if fusion_layers == "all":                                  # deepavfusion.py:38-46
            fusion_layers = set(range(max_depth))
        elif fusion_layers == "none":
            fusion_layers = set()
        elif isinstance(fusion_layers, int):
            fusion_layers = {fusion_layers}
        else:
            fusion_layers = {int(l) for l in str(fusion_layers).split("-")}
        self.fusion_blocks = nn.ModuleList([
            None if i not in fusion_layers else FusionBlock(
                dim=self.embed_dim, num_heads=fusion_num_heads, attn_ratio=fusion_attn_ratio, mlp_ratio=fusion_mlp_ratio,
                qkv_bias=True, drop=drop, attn_drop=attn_drop, drop_path=drop_path, norm_layer=nn.LayerNorm)
            for i in range(max_depth)])
        self.fusion_norm = FinalNorm(self.embed_dim)
        self.initialize_weights()

    def initialize_weights(self):
        nn.init.normal_(self.fusion_tokens, std=0.02)
        self.fusion_blocks.apply(_xavier_linear_)

    def params_layer_ids(self):
        ids = []
        ids.extend(self.image.params_layer_ids())
        ids.extend(self.audio.params_layer_ids())
        ids.append((self.fusion_tokens, 0))
        for i, blk in enumerate(self.fusion_blocks):
            if blk is not None:
                ids.extend([(p, i + 1) for p in blk.parameters()])
        ids.extend([(p, len(self.fusion_blocks) + 1) for p in self.fusion_norm.parameters()])
        return ids

    def load_checkpoint(self, ckpt_fn, prefix):
        ckpt = torch.load(ckpt_fn, map_location="cpu")["state_dict"]
        ckpt = {k[len(prefix):]: ckpt[k] for k in ckpt if k.startswith(prefix)}
        self.load_state_dict(ckpt, strict=True)
        print(f"Loaded pre-trained checkpoint: {ckpt_fn}")

    def forward(self, image, audio, image_ids_keep=None, audio_ids_keep=None, return_embs=False):
        with ensure_store(self):
            return self._forward(image, audio, image_ids_keep, audio_ids_keep, return_embs)

    def _bind(self, store):
        self._tok_ns = SimpleNamespace(store=store, tokens=self.fusion_tokens)

    def _side_streams(self, device):
        """Two side streams: within a layer the image block, the audio block and the fusion block read the
        same inputs and are independent, so they are issued on three streams.  Autograd runs
        each node's backward on its forward stream, so backward overlaps the same way, and a CUDA-graph
        capture of the step records the branches as parallel graph paths.  The many small fusion-block
        kernels (grids of 8-72 CTAs) then fill SMs the modality blocks leave idle."""
        if device.type != "cuda" or os.environ.get("DAVF_STREAMS", "1") == "0":
            return None
        st = self.__dict__.get("_davf_streams")
        if st is None or st[0].device != device:
            # The fusion block is the longest dependent chain of a layer (many small kernels): its stream gets the
            # higher priority, so its CTAs are dispatched first whenever a persistent GEMM of another branch retires.
            prio = int(os.environ.get("DAVF_FUSION_PRIORITY", "-1"))
            st = (torch.cuda.Stream(device), torch.cuda.Stream(device, priority=prio))
            self.__dict__["_davf_streams"] = st
        store = self.__dict__.get("_davf_store")
        if store is not None:
            if st[0] not in store.side_streams:
                store.side_streams.extend(st)
            store.main_stream = torch.cuda.current_stream()
        return st

    def _forward(self, image, audio, image_ids_keep, audio_ids_keep, return_embs):
        B = image.shape[0]
        side = self._side_streams(image.device)
        cur = torch.cuda.current_stream() if side else None

        def on(stream):
            return torch.cuda.stream(stream) if side else contextlib.nullcontext()

        def fork():
            if side:
                side[0].wait_stream(cur)
                side[1].wait_stream(cur)

        def join():
            if side:
                cur.wait_stream(side[0])
                cur.wait_stream(side[1])

        def share(t, *streams):
            """``t`` is read by kernels on ``streams`` other than the one it was allocated on: tell the caching
            allocator, or its block could be re-used (by the allocating stream) while those kernels still run
            -- in particular when autograd frees saved activations during the multi-stream backward."""
            if side:
                for s in streams:
                    t.record_stream(s)
            return t

        fork()
        x_image = self.image.prepare_patch_tokens(image, image_ids_keep)      # (B, nI, D) f32
        with on(side[0] if side else None):
            x_audio = self.audio.prepare_patch_tokens(audio, audio_ids_keep)  # (B, nA, D)
        join()
        embs = []
        x_fusion = Fn.BroadcastTokensFn.apply(self.fusion_tokens, B, self._tok_ns)
        if side:
            share(x_image, side[1]); share(x_audio, cur, side[1]); share(x_fusion, side[0], side[1])
        for blk_image, blk_audio, blk_fusion in zip(self.image.blocks, self.audio.blocks, self.fusion_blocks):
            fork()
            if blk_fusion is None:
                x_image = blk_image(x_image)
                with on(side[0] if side else None):
                    x_audio = blk_audio(x_audio)
            else:
                # deepavfusion.py:104-107: the modality blocks see the fusion tokens as extra keys / values;
                # the fusion block reads the PRE-block modality tokens.  Each input has two / three consumers: FanOutFn
                # makes the gradient fan-in one kernel, run on the stream of the branch that consumes the sum.
                store = self.__dict__["_davf_store"]
                xi_blk, xi_fus = Fn.FanOutFn.apply(x_image, 2, store)
                with on(side[0] if side else None):
                    xa_blk, xa_fus = Fn.FanOutFn.apply(x_audio, 2, store)
                with on(side[1] if side else None):
                    xf_img, xf_aud, xf_fus = Fn.FanOutFn.apply(x_fusion, 3, store)
                _x_image = blk_image(xi_blk, prefix=xf_img)
                with on(side[0] if side else None):
                    _x_audio = blk_audio(xa_blk, prefix=xf_aud)
                with on(side[1] if side else None):
                    x_fusion = blk_fusion(xf_fus, xi_fus, xa_fus)
                x_image, x_audio = _x_image, _x_audio
            join()
            if side:      # next layer (and the final norms on `cur`) read these across streams
                share(x_image, side[1]); share(x_audio, cur, side[1]); share(x_fusion, cur, side[0])
            if return_embs:
                embs.append((x_image, x_audio, x_fusion))
        x_image = self.image.norm(x_image)
        x_audio = self.audio.norm(x_audio)
        x_fusion = self.fusion_norm(x_fusion)
        if not return_embs:
            return x_image, x_audio, x_fusion
        return x_image, x_audio, x_fusion, embs

"""Audio-visual masked auto-encoder objective (mirror of reference models/avmae.py:9-236).

Drop-in: ``AVMAE(encoder, encoder_dim, image_decoder_arch, image_decoder_depth, image_mask_ratio,
image_norm_loss, audio_..., decoder_dim=512, num_heads=16, mlp_ratio=4.)``,
``forward(image, audio) -> (loss_image, loss_audio, pred_image, pred_audio)``, ``forward_encoder``,
``random_masking``, ``patchify`` and the ``state_dict`` layout are the reference's.  Only the ``plain``
decoder is built (``decoder_arch: plain`` in every config; the Swin decoder is out of scope).
"""
from __future__ import annotations

import contextlib
from types import SimpleNamespace

import torch
from torch import nn

from .. import functional as Fn
from .. import kernels as K
from ..util.pos_embed import get_2d_sincos_pos_embed
from .layers import Block, ensure_store


class AVMAE(nn.Module):
    def __init__(self, encoder, encoder_dim,
                 image_decoder_arch="plain", image_decoder_depth=8, image_mask_ratio=0.75, image_norm_loss=False,
                 audio_decoder_arch="plain", audio_decoder_depth=8, audio_mask_ratio=0.8, audio_norm_loss=False,
                 decoder_dim=512, num_heads=16, mlp_ratio=4.0, norm_layer=nn.LayerNorm):
        super().__init__()
        if image_decoder_arch != "plain" or audio_decoder_arch != "plain":
            raise NotImplementedError("only decoder_arch='plain' (configs/deepavfusion.yaml:16,23) is built")
        self.image_mask_ratio, self.image_norm_loss = image_mask_ratio, image_norm_loss
        self.audio_mask_ratio, self.audio_norm_loss = audio_mask_ratio, audio_norm_loss
        self.decoder_dim = decoder_dim
        self.encoder = encoder
        self.image_gs, self.audio_gs = encoder.image.patch_embed.grid_size, encoder.audio.patch_embed.grid_size
        self.image_ps, self.audio_ps = encoder.image.patch_embed.patch_size, encoder.audio.patch_embed.patch_size

        # registration order follows avmae.py:31-89 (audio decoder first) so state_dict order matches
        for mod, gs, ps, depth, chans in (("audio", self.audio_gs, self.audio_ps, audio_decoder_depth, 1),
                                          ("image", self.image_gs, self.image_ps, image_decoder_depth, 3)):
            setattr(self, f"{mod}_decoder_embed", nn.Linear(encoder_dim, decoder_dim, bias=True))
            setattr(self, f"{mod}_decoder_mask_token", nn.Parameter(torch.zeros(1, 1, decoder_dim)))
            setattr(self, f"{mod}_decoder_pos_embed", nn.Parameter(torch.zeros(1, gs[0] * gs[1], decoder_dim)))
            setattr(self, f"{mod}_decoder_arch", "plain")
            setattr(self, f"{mod}_decoder_blocks", nn.ModuleList([
                Block(decoder_dim, num_heads, mlp_ratio, qkv_bias=True, norm_layer=norm_layer) for _ in range(depth)]))
            setattr(self, f"{mod}_decoder_norm", norm_layer(decoder_dim))
            setattr(self, f"{mod}_decoder_pred", nn.Linear(decoder_dim, ps[0] * ps[1] * chans, bias=True))
        self.initialize_weights()

    def initialize_weights(self):
        """avmae.py:92-107: sin-cos decoder pos-embeds (trainable), N(0,.02) mask tokens, xavier decoders."""
        for mod, gs in (("image", self.image_gs), ("audio", self.audio_gs)):
            pe = get_2d_sincos_pos_embed(self.decoder_dim, gs, cls_token=False)
            getattr(self, f"{mod}_decoder_pos_embed").data.copy_(torch.from_numpy(pe).float().unsqueeze(0))
            nn.init.normal_(getattr(self, f"{mod}_decoder_mask_token"), std=0.02)
        for n, m in self.named_modules():
            if not n.startswith("encoder"):
                self._init_weights(m)

    @staticmethod
    def _init_weights(m):
        if isinstance(m, nn.Linear):
            nn.init.xavier_uniform_(m.weight)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    # ------------------------------------------------------------------------------------------
    def _bind(self, store):
        self._dec = {}
        for mod, ratio, norm_loss, gs, ps in (("image", self.image_mask_ratio, self.image_norm_loss, self.image_gs, self.image_ps),
                                              ("audio", self.audio_mask_ratio, self.audio_norm_loss, self.audio_gs, self.audio_ps)):
            embed, norm, pred = (getattr(self, f"{mod}_decoder_{k}") for k in ("embed", "norm", "pred"))
            L = gs[0] * gs[1]
            e = SimpleNamespace(store=store, embed_w=embed.weight, embed_b=embed.bias,
                                mask_token=getattr(self, f"{mod}_decoder_mask_token"),
                                pos_embed=getattr(self, f"{mod}_decoder_pos_embed"))
            h = SimpleNamespace(store=store, eps=norm.eps, norm_w=norm.weight, norm_b=norm.bias,
                                pred_w=pred.weight, pred_b=pred.bias, patch=ps[0], norm_pix=bool(norm_loss),
                                len_keep=int(L * (1 - ratio)))
            self._dec[mod] = (e, h)

    def random_masking(self, N, L, mask_ratio, device):
        """avmae.py:120-142.  The noise is still ``torch.rand(N, L)`` on the device (same RNG stream as
        the reference); the two argsorts + gathers are one rank-counting kernel.  Ties (P ~ 1e-3 per
        row at L=196) are broken lower-index-first, i.e. like ``argsort(stable=True)``."""
        noise = torch.rand(N, L, device=device)
        len_keep = int(L * (1 - mask_ratio))
        ids_restore, ids_keep, mask = K.mask_rank(noise, len_keep)
        return ids_keep, mask, ids_restore

    def forward_encoder(self, image, audio):
        return self.encoder(image, audio)

    def forward_decoder(self, x, x_fusion, ids_restore, modality="image", ids_keep=None):
        """avmae.py:147-180 up to (and excluding) the final norm + pred, which are fused with the loss.
        Returns the decoder token sequence [B, nFus + L, decoder_dim]."""
        e, _ = self._dec[modality]
        if ids_keep is None:        # inverse of ids_restore restricted to the kept ranks
            nK = x.shape[1]
            ids_keep = torch.argsort(ids_restore, dim=1)[:, :nK].contiguous()
        seq = Fn.DecoderEmbedFn.apply(x, x_fusion, ids_restore, ids_keep, e.embed_w, e)
        for blk in getattr(self, f"{modality}_decoder_blocks"):
            seq = blk(seq)
        return seq

    @staticmethod
    def patchify(x, patch_size):
        """avmae.py:201-214 (host-visible helper; the loss kernel reads NCHW directly and never calls it)."""
        bs, c = x.shape[:2]
        pH, pW = patch_size
        gH, gW = x.shape[2] // pH, x.shape[3] // pW
        x = x.reshape(bs, c, gH, pH, gW, pW)
        x = torch.einsum("nchpwq->nhwpqc", x)
        return x.reshape(bs, gH * gW, pH * pW * c)

    def forward(self, image, audio):
        with ensure_store(self):
            return self._forward(image, audio)

    def _forward(self, image, audio):
        B, device = image.shape[0], image.device
        image = image.float().contiguous()
        audio = audio.float().contiguous()
        Li, La = self.image_gs[0] * self.image_gs[1], self.audio_gs[0] * self.audio_gs[1]
        # two torch.rand draws, image first (avmae.py:220-221)
        image_ids_keep, image_mask, image_ids_restore = self.random_masking(B, Li, self.image_mask_ratio, device)
        audio_ids_keep, audio_mask, audio_ids_restore = self.random_masking(B, La, self.audio_mask_ratio, device)

        x_image, x_audio, x_fusion = self.encoder(image, audio, image_ids_keep=image_ids_keep, audio_ids_keep=audio_ids_keep)

        # the two decoders are independent: issue the audio decoder on a side stream (see DeepAVFusion._side_streams)
        side = self.encoder._side_streams(device)
        cur = torch.cuda.current_stream() if side else None
        if side:
            side[0].wait_stream(cur)
            for t in (x_audio, x_fusion, audio, audio_mask, audio_ids_restore, audio_ids_keep):
                t.record_stream(side[0])              # allocated on `cur`, read by the audio decoder's stream
        seq_i = self.forward_decoder(x_image, x_fusion, image_ids_restore, "image", image_ids_keep)
        hi = self._dec["image"][1]
        loss_image, pred_image = Fn.PredLossFn.apply(seq_i, image, image_mask, hi.norm_w, hi)

        with (torch.cuda.stream(side[0]) if side else contextlib.nullcontext()):
            seq_a = self.forward_decoder(x_audio, x_fusion, audio_ids_restore, "audio", audio_ids_keep)
            ha = self._dec["audio"][1]
            loss_audio, pred_audio = Fn.PredLossFn.apply(seq_a, audio, audio_mask, ha.norm_w, ha)
        if side:
            cur.wait_stream(side[0])
            loss_audio.record_stream(cur)
            pred_audio.record_stream(cur)
        return loss_image, loss_audio, pred_image, pred_audio

"""Audio-visual classifier (mirror of reference models/classifier.py:4-64), BASELINE configs 4-5:
the unmasked DeepAVFusion encoder, token mean-pool, optional BatchNorm1d(affine=False) and three
Linear heads.  Same constructor, ``forward(image, audio) -> (pred_image, pred_audio, pred_fusion)``,
``params_layer_ids()``, ``train()`` behaviour and ``state_dict`` keys (``encoder.*``,
``{image,audio,fusion}_norm.{running_mean,running_var,num_batches_tracked}``, ``*_head.{weight,bias}``).

The encoder is the tensor-core path; the tail (pool / BN / heads) is three tiny f32 kernels per
modality (``functional.ClassifierTailFn``) -- the reference runs lin-probe and fine-tuning with
``use_amp: False`` (configs/linprobe.yaml:35, configs/finetune.yaml:50).
"""
from __future__ import annotations

from types import SimpleNamespace

import torch
from torch import nn

from .. import functional as Fn
from .layers import ensure_store


class AVClassifier(nn.Module):
    def __init__(self, encoder, num_classes, freeze_encoder=False, input_norm=False):
        super().__init__()
        self.encoder = encoder
        self.freeze_encoder = freeze_encoder
        if self.freeze_encoder:
            for p in self.encoder.parameters():
                p.requires_grad = False
        self.input_norm = input_norm
        if self.input_norm:                                           # classifier.py:14-18
            self.image_norm = nn.BatchNorm1d(self.encoder.embed_dim, affine=False, eps=1e-6)
            self.audio_norm = nn.BatchNorm1d(self.encoder.embed_dim, affine=False, eps=1e-6)
            self.fusion_norm = nn.BatchNorm1d(self.encoder.embed_dim, affine=False, eps=1e-6)
        self.image_head = nn.Linear(self.encoder.embed_dim, num_classes)
        self.audio_head = nn.Linear(self.encoder.embed_dim, num_classes)
        self.fusion_head = nn.Linear(self.encoder.embed_dim, num_classes)
        self.initialize_weights()

    def initialize_weights(self):
        for head in (self.image_head, self.audio_head, self.fusion_head):
            nn.init.xavier_uniform_(head.weight)
            nn.init.zeros_(head.bias)

    def params_layer_ids(self):
        ids = []
        ids.extend(self.encoder.params_layer_ids())
        top = len(self.encoder.audio.blocks) + 1
        for head in (self.image_head, self.audio_head, self.fusion_head):
            ids.extend([(p, top) for p in head.parameters()])
        return ids

    def _bind(self, store):
        def ns(head, bn):
            return SimpleNamespace(store=store, head_w=head.weight, head_b=head.bias, bn=bn)
        self._tails = (ns(self.image_head, self.image_norm if self.input_norm else None),
                       ns(self.audio_head, self.audio_norm if self.input_norm else None),
                       ns(self.fusion_head, self.fusion_norm if self.input_norm else None))

    def forward(self, image, audio):
        with ensure_store(self):
            image = image.float().contiguous()
            audio = audio.float().contiguous()
            if self.freeze_encoder:
                with torch.no_grad():
                    x_image, x_audio, x_fusion = self.encoder(image, audio)
            else:
                x_image, x_audio, x_fusion = self.encoder(image, audio)
            ti, ta, tf = self._tails
            # BatchNorm1d uses batch statistics iff the module is in training mode (classifier.py:50-54)
            pred_image = Fn.ClassifierTailFn.apply(x_image, ti.head_w, ti, self.input_norm and self.image_norm.training)
            pred_audio = Fn.ClassifierTailFn.apply(x_audio, ta.head_w, ta, self.input_norm and self.audio_norm.training)
            pred_fusion = Fn.ClassifierTailFn.apply(x_fusion, tf.head_w, tf, self.input_norm and self.fusion_norm.training)
            return pred_image, pred_audio, pred_fusion

    def train(self, mode: bool = True):
        super().train(mode)
        if self.freeze_encoder:
            self.encoder.train(False)
        # classifier.py:61-64 returns None here (reference quirk); kept

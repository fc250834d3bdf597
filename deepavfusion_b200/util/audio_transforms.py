"""GPU input stage (mirror of the transforms the reference composes on CPU workers: util/audio_transforms.py and
train.py:44-54).  ``GpuLogMel`` = [RandomVol ->] MelSpectrogram(sample_rate, n_fft, hop_length, n_mels) -> Log, evaluated by
one sm_100a kernel on a batch of raw waveforms (f32 or int16 PCM) that were copied to the device as they are;
``GpuNormalize`` = ToTensor + Normalize on uint8 frames.  The data loader then ships 2-byte samples and 1-byte pixels over
PCIe and the model receives exactly the tensors ``train.py:159-160`` would have handed it."""
from __future__ import annotations

from typing import Optional, Sequence

import torch

from .. import kernels as K


class GpuLogMel:
    def __init__(self, sample_rate: int = 16000, n_fft: int = 800, hop_length: int = 250, n_mels: int = 128, eps: float = 1e-7,
                 drop_last_frame: bool = True):
        """``drop_last_frame``: the ``[:, :, :-1]`` of datasets.py:242 (3 s at 16 kHz -> 192 frames)."""
        self.sample_rate, self.n_fft, self.hop, self.n_mels, self.eps, self.drop = sample_rate, n_fft, hop_length, n_mels, eps, drop_last_frame
        self._ws = {}

    def __call__(self, waveform: torch.Tensor, gain_db: Optional[torch.Tensor] = None) -> torch.Tensor:
        """waveform [B, T] (f32 in [-1, 1] or int16 PCM) on the GPU -> [B, 1, n_mels, T // hop (+1)] f32.
        ``gain_db`` [B]: RandomVol's per-clip gain (drawn by the caller, uniform(-6, 6) in the reference)."""
        dev = waveform.device
        ws = self._ws.get(dev)
        if ws is None:
            ws = self._ws[dev] = K.logmel_workspace(self.sample_rate, self.n_fft, self.hop, self.n_mels, dev)
        frames = waveform.shape[1] // self.hop + (0 if self.drop else 1)
        return K.logmel_fwd(ws, waveform.contiguous(), gain_db, self.n_mels, frames, self.eps)


class GpuNormalize:
    def __init__(self, mean: Sequence[float] = (0.485, 0.456, 0.406), std: Sequence[float] = (0.229, 0.224, 0.225)):
        self.mean, self.std = tuple(mean), tuple(std)

    def __call__(self, frames_u8: torch.Tensor) -> torch.Tensor:
        """uint8 [B, H, W, C] on the GPU -> f32 [B, C, H, W]."""
        return K.image_normalize_u8(frames_u8.contiguous(), self.mean, self.std)

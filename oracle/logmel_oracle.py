"""CPU oracle of the audio front-end that sits just in front of the hot path.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Restates, in plain torch ops on CPU, what the reference builds in
train.py:50-54 from ``torchaudio.transforms`` (util/audio_transforms.py:3) -- ``RandomVol`` (gain in dB, clamp to [-1, 1];
audio_transforms.py:8-18), ``MelSpectrogram(sample_rate=16000, n_fft=800, hop_length=250, n_mels=128)`` (torchaudio
defaults: periodic Hann window of n_fft samples, centre-padded with reflection, power 2, one-sided, HTK mel scale, no
filter normalisation, f_min 0, f_max sr / 2), ``Log`` = log10(x + 1e-7) (audio_transforms.py:30-36) -- and the
``[..., :-1]`` frame crop of datasets.py:242.  ``oracle/make_golden_logmel.py`` pins it against the installed torchaudio
and writes tests/golden/logmel.npz.
"""
from __future__ import annotations

import math

import torch


def hann_periodic(n: int) -> torch.Tensor:
    k = torch.arange(n, dtype=torch.float64)
    return (0.5 - 0.5 * torch.cos(2.0 * math.pi * k / n))


def mel_filterbank_htk(n_freqs: int, f_min: float, f_max: float, n_mels: int, sample_rate: int) -> torch.Tensor:
    """torchaudio.functional.melscale_fbanks(norm=None, mel_scale='htk'): [n_freqs, n_mels] triangular filters."""
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs, dtype=torch.float64)
    m_min = 2595.0 * math.log10(1.0 + f_min / 700.0)
    m_max = 2595.0 * math.log10(1.0 + f_max / 700.0)
    m_pts = torch.linspace(m_min, m_max, n_mels + 2, dtype=torch.float64)
    f_pts = 700.0 * (10.0 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down = -slopes[:, :-2] / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return torch.clamp(torch.minimum(down, up), min=0.0)


def log_mel(waveform: torch.Tensor, gain_db: torch.Tensor | None = None, sample_rate: int = 16000, n_fft: int = 800,
            hop: int = 250, n_mels: int = 128, eps: float = 1e-7, drop_last: bool = True) -> torch.Tensor:
    """waveform f32 [B, T] in [-1, 1] -> log-mel f32 [B, 1, n_mels, T // hop (+1 without drop_last)]."""
    x = waveform.to(torch.float64)
    if gain_db is not None:                                   # RandomVol: F.gain + clamp
        x = torch.clamp(x * (10.0 ** (gain_db.to(torch.float64) / 20.0)).unsqueeze(1), -1.0, 1.0)
    xp = torch.nn.functional.pad(x.unsqueeze(1), (n_fft // 2, n_fft // 2), mode="reflect").squeeze(1)
    frames = xp.unfold(1, n_fft, hop) * hann_periodic(n_fft)                  # [B, F, n_fft]
    spec = torch.fft.rfft(frames, dim=-1)
    power = spec.real ** 2 + spec.imag ** 2                                    # [B, F, n_fft // 2 + 1]
    mel = power @ mel_filterbank_htk(n_fft // 2 + 1, 0.0, sample_rate / 2.0, n_mels, sample_rate)
    out = torch.log10(mel + eps).transpose(1, 2).unsqueeze(1)                  # [B, 1, n_mels, F]
    if drop_last:
        out = out[..., :-1]
    return out.to(torch.float32)


def make_wave(B=3, T=48000, seed=0):
    """Speech-like test signal: a few drifting tones + coloured noise at very different levels (70 dB of dynamic range)."""
    g = torch.Generator().manual_seed(seed)
    t = torch.arange(T, dtype=torch.float64) / 16000.0
    w = torch.zeros(B, T, dtype=torch.float64)
    for b in range(B):
        for _ in range(4):
            f0, a = 80.0 + 3000.0 * torch.rand((), generator=g).item(), 10.0 ** (-3.0 * torch.rand((), generator=g).item())
            w[b] += a * torch.sin(2 * math.pi * (f0 * t + 40.0 * t * t))
        noise = torch.randn(T, generator=g, dtype=torch.float64)
        w[b] += 1e-3 * torch.cumsum(noise, 0) / 30.0 + 1e-4 * noise
    return (w / w.abs().max() * 0.9).to(torch.float32)


def normalize_image_u8(img_u8: torch.Tensor, mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225)) -> torch.Tensor:
    """vT.ToTensor() + vT.Normalize (train.py:48-49) on a uint8 [B, H, W, 3] batch -> f32 [B, 3, H, W]."""
    x = img_u8.to(torch.float32).permute(0, 3, 1, 2) / 255.0
    m = torch.tensor(mean, dtype=torch.float32).view(1, 3, 1, 1)
    s = torch.tensor(std, dtype=torch.float32).view(1, 3, 1, 1)
    return (x - m) / s

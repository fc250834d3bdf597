"""-m gpu: the GPU input stage against its CPU oracle and against the fixture written from the REAL
torchaudio pipeline of the reference (oracle/make_golden_logmel.py).  Tolerance: the output is log10 of a power that spans
> 100 dB; all arithmetic is f32, the quietest mel bands sit on the f32 rounding floor of the transform (torchaudio's own f32
result differs from an f64 evaluation by 8e-4 there): abs 3e-3 everywhere, 3e-4 on the bands within 40 dB of their frame's
loudest band."""
import os

import numpy as np
import pytest
import torch

from oracle import logmel_oracle as L

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _check(out, ref):
    err = (out - ref).abs()
    assert float(err.max()) < 3e-3, float(err.max())
    loud = ref > ref.amax(dim=2, keepdim=True) - 4.0
    assert float(err[loud].max()) < 3e-4, float(err[loud].max())


def test_logmel_golden_and_oracle():
    from deepavfusion_b200.util.audio_transforms import GpuLogMel
    z = np.load(os.path.join(GOLD, "logmel.npz"))
    wave = L.make_wave()
    assert np.array_equal(wave[:, :64].numpy(), z["wave_head"])          # the recipe regenerates the fixture's input
    gain = torch.from_numpy(z["gain_db"])
    fe = GpuLogMel()
    out = fe(wave.cuda(), gain.cuda()).cpu()
    assert out.shape == (3, 1, 128, 192)
    _check(out, torch.from_numpy(z["logmel"]))                            # the reference's torchaudio pipeline
    _check(out, L.log_mel(wave, gain))                                    # the oracle
    # int16 PCM input (what crosses PCIe), no gain, all 193 frames
    pcm = torch.round(wave * 32767.0).to(torch.int16)
    out16 = GpuLogMel(drop_last_frame=False)(pcm.cuda()).cpu()
    assert out16.shape == (3, 1, 128, 193)
    _check(out16, L.log_mel(pcm.float() / 32768.0, None, drop_last=False))


@pytest.mark.parametrize("B,T,n_mels", [(64, 48000, 128), (1, 16000, 64), (5, 1001, 80)])
def test_logmel_shapes(B, T, n_mels):
    from deepavfusion_b200.util.audio_transforms import GpuLogMel
    g = torch.Generator().manual_seed(B)
    wave = (torch.randn(B, T, generator=g) * 0.1).clamp(-1, 1)
    out = GpuLogMel(n_mels=n_mels)(wave.cuda()).cpu()
    ref = L.log_mel(wave, None, n_mels=n_mels)
    assert out.shape == ref.shape == (B, 1, n_mels, T // 250)
    _check(out, ref)


def test_image_normalize():
    from deepavfusion_b200.util.audio_transforms import GpuNormalize
    img = torch.randint(0, 256, (7, 224, 224, 3), dtype=torch.uint8, generator=torch.Generator().manual_seed(3))
    out = GpuNormalize()(img.cuda()).cpu()
    ref = L.normalize_image_u8(img)
    assert out.shape == (7, 3, 224, 224) and float((out - ref).abs().max()) < 1e-5

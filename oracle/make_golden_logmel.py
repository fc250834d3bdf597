"""Pins oracle/logmel_oracle.py against the reference's own front-end -- torchaudio.transforms.MelSpectrogram with the
arguments of train.py:53 followed by util/audio_transforms.Log and the [..., :-1] crop of datasets.py:242 -- and writes
tests/golden/logmel.npz (seeded waveform recipe + expected log-mel of the REAL torchaudio pipeline).

    python oracle/make_golden_logmel.py        # needs torchaudio (present in the build container)
"""
import os
import sys

import numpy as np
import torch
import torchaudio

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import logmel_oracle as L  # noqa: E402


def main():
    wave = L.make_wave()
    gain = torch.tensor([0.0, -6.0, 5.0])
    # the reference pipeline (train.py:50-54): RandomVol (with the gain fixed), MelSpectrogram, Log; datasets.py:242 crop
    mel = torchaudio.transforms.MelSpectrogram(sample_rate=16000, n_fft=800, hop_length=250, n_mels=128)
    ref = []
    for b in range(wave.shape[0]):
        x = torch.clamp(torchaudio.functional.gain(wave[b:b + 1], float(gain[b])), -1, 1)
        ref.append(torch.log10(mel(x) + 1e-7)[:, :, :-1])
    ref = torch.stack(ref)                                        # [B, 1, 128, 192]
    ours = L.log_mel(wave, gain)
    err = (ours - ref).abs().max().item()
    print("oracle vs torchaudio: max abs err", err, "shape", tuple(ref.shape))
    # torchaudio computes in f32: bins 70 dB below the loudest component sit on its FFT rounding floor (a few 1e-4 in log10)
    assert ours.shape == ref.shape == (3, 1, 128, 192) and err < 2e-3, err
    fb = torchaudio.functional.melscale_fbanks(401, 0.0, 8000.0, 128, 16000, norm=None, mel_scale="htk")
    assert (L.mel_filterbank_htk(401, 0.0, 8000.0, 128, 16000).float() - fb).abs().max().item() < 1e-4      # (torchaudio builds the filters in f32)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "logmel.npz"), gain_db=gain.numpy(), logmel=ref.numpy(),
                        wave_head=wave[:, :64].numpy())
    print("wrote tests/golden/logmel.npz")


if __name__ == "__main__":
    main()

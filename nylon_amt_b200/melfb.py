"""Mel filterbank / window tables for the fused log-mel kernel, built with the same fp32 torch ops torchaudio uses
(torchaudio.functional.melscale_fbanks(mel_scale='htk', norm='slaney') and torch.hann_window), so the tables are
bit-identical to the ones the reference's MelSpectrogram holds (hftt_code/model/amt.py:59) without importing
torchaudio.  The arithmetic order matters: f_pts - all_freqs cancels, so a 1-ulp difference in pow shows up at 1e-5
relative in the weights (tests/test_host_logic.py pins the result against tests/golden/mel_fb.npz).
"""
import math

import torch


def melscale_fbanks(n_freqs=1025, f_min=0.0, f_max=8000.0, n_mels=256, sample_rate=16000):
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    m_min = 2595.0 * math.log10(1.0 + (f_min / 700.0))
    m_max = 2595.0 * math.log10(1.0 + (f_max / 700.0))
    m_pts = torch.linspace(m_min, m_max, n_mels + 2)
    f_pts = 700.0 * (10.0 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    zero = torch.zeros(1)
    down_slopes = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up_slopes = slopes[:, 2:] / f_diff[1:]
    fb = torch.max(zero, torch.min(down_slopes, up_slopes))
    enorm = 2.0 / (f_pts[2:n_mels + 2] - f_pts[:n_mels])
    fb = fb * enorm.unsqueeze(0)
    return fb.contiguous()


def hann_window(n_fft=2048):
    return torch.hann_window(n_fft, periodic=True, dtype=torch.float32)

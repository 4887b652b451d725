"""CPU restatement of amt.py:56-58 (channel mean + torchaudio.transforms.Resample) -- TEST INFRASTRUCTURE.
The arithmetic lives in torchaudio (functional.py `_get_sinc_resample_kernel` / `_apply_sinc_resample_kernel`, unpinned by the
reference; pinned here against torchaudio 2.11.0 through tests/golden/resample.npz, written by oracle/make_golden_resample.py
from the unmodified reference)."""
import math

import numpy as np


def sinc_kernel(orig, new, lowpass_filter_width=6, rolloff=0.99):
    """The polyphase table of sinc_interp_hann: [new/g][2*width + orig/g] (float64 math, float32 result) and width."""
    g = math.gcd(int(orig), int(new))
    o, n = int(orig) // g, int(new) // g
    base = min(o, n) * rolloff
    width = math.ceil(lowpass_filter_width * o / base)
    idx = np.arange(-width, width + o, dtype=np.float64)[None, :] / o
    # torchaudio divides an int64 arange by new_freq with dtype=None: the phase offsets -p/n are rounded to float32 before they
    # meet the float64 index grid (functional.py: `torch.arange(0, -new_freq, -1, dtype=dtype)[:, None, None] / new_freq + idx`)
    t = (np.arange(0, -n, -1).astype(np.float32) / np.float32(n)).astype(np.float64)[:, None] + idx
    t = np.clip(t * base, -lowpass_filter_width, lowpass_filter_width)
    window = np.cos(t * math.pi / lowpass_filter_width / 2) ** 2
    t = t * math.pi
    scale = base / o
    k = np.where(t == 0, 1.0, np.sin(t) / np.where(t == 0, 1.0, t)) * window * scale
    return k.astype(np.float32), width, o, n


def mono_resample(wave, orig, new=16000):
    """wave float32 [C, N] -> float32 [ceil(new*N/orig)]: mean over channels, zero-pad (width, width + o), strided FIR per phase."""
    x = np.mean(np.asarray(wave, dtype=np.float32), axis=0, dtype=np.float32)
    if int(orig) == int(new):
        return x
    k, width, o, n = sinc_kernel(orig, new)
    N = x.shape[0]
    xp = np.concatenate([np.zeros(width, np.float32), x, np.zeros(width + o, np.float32)])
    kw = k.shape[1]
    n_i = (xp.shape[0] - kw) // o + 1
    win = np.lib.stride_tricks.as_strided(xp, (n_i, kw), (xp.strides[0] * o, xp.strides[0]))
    y = (win.astype(np.float64) @ k.T.astype(np.float64)).astype(np.float32).reshape(-1)      # [n_i, n] -> j = i*n + p
    return y[: int(math.ceil(n * N / o))]

"""Writes tests/golden/resample.npz (build container only): non-16 kHz / multi-channel clips through the UNMODIFIED
reference AMT.wav2feature (hftt_code/model/amt.py:55-61), plus the intermediate mono 16 kHz waveform computed with the
same two torchaudio calls amt.py:56-58 makes.

    python -m oracle.make_golden_resample
"""
import os
import sys
import tempfile
import wave as _wave

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import _refload  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def write_wav(path, pcm, sr):                      # pcm int16 [C, N]
    with _wave.open(path, "wb") as f:
        f.setnchannels(pcm.shape[0])
        f.setsampwidth(2)
        f.setframerate(sr)
        f.writeframes(np.ascontiguousarray(pcm.T).tobytes())


def main():
    import torchaudio
    ref_amt, _ = _refload.load()
    cfg = _refload.config()
    A = ref_amt.AMT(cfg, None, None)
    rng = np.random.default_rng(5)
    d = {}
    cases = [("s44100c2", 44100, 2, 44100 * 2 + 17), ("s48000c1", 48000, 1, 48000 + 5), ("s22050c2", 22050, 2, 30011), ("s8000c1", 8000, 1, 12001),
             ("s16000c2", 16000, 2, 20000)]
    for name, sr, C, N in cases:
        t = np.arange(N) / sr
        x = 0.2 * np.sin(2 * np.pi * 440.0 * t)[None] + 0.1 * np.sin(2 * np.pi * 3000.0 * t)[None] + 0.05 * rng.standard_normal((C, N))
        pcm = np.clip(np.round(x * 32768.0), -32768, 32767).astype("<i2")
        with tempfile.TemporaryDirectory() as tmp:
            p = os.path.join(tmp, name + ".wav")
            write_wav(p, pcm, sr)
            feat = A.wav2feature(p)
        wave = torch.from_numpy(pcm.astype(np.float32) / np.float32(32768.0))
        mono16k = torchaudio.transforms.Resample(sr, cfg["feature"]["sr"])(torch.mean(wave, dim=0))
        d["pcm_" + name] = pcm
        d["sr_" + name] = np.int32(sr)
        d["mono16k_" + name] = mono16k.numpy()
        d["feat_" + name] = feat.numpy()
        print(name, pcm.shape, mono16k.shape, tuple(feat.shape))
    np.savez_compressed(os.path.join(OUT, "resample.npz"), **d)


if __name__ == "__main__":
    main()

"""Channel mix-down + resampling in front of the log-mel kernel (amt.py:56-58): oracle vs the reference-written fixture (CPU),
CUDA kernel vs the fixture and the oracle (GPU)."""
import os
import wave as _wave

import numpy as np
import pytest
import torch

from oracle import logmel_oracle, resample_oracle

CASES = ["s44100c2", "s48000c1", "s22050c2", "s8000c1", "s16000c2"]


@pytest.fixture(scope="module")
def fx(golden_dir):
    return np.load(os.path.join(golden_dir, "resample.npz"))


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference(fx, name):
    wave = fx["pcm_" + name].astype(np.float32) / np.float32(32768.0)
    y = resample_oracle.mono_resample(wave, int(fx["sr_" + name]))
    ref = fx["mono16k_" + name]
    assert y.shape == ref.shape
    assert float(np.abs(y - ref).max()) <= 2e-6, name


def test_oracle_table_equals_torchaudio():
    from torchaudio.functional.functional import _get_sinc_resample_kernel
    for sr in (44100, 48000, 22050, 8000):
        import math
        g = math.gcd(sr, 16000)
        k, w = _get_sinc_resample_kernel(sr, 16000, g)
        mine, width, o, n = resample_oracle.sinc_kernel(sr, 16000)
        assert width == w and mine.shape == (n, 2 * width + o)
        assert float(np.abs(mine - k.reshape(n, -1).numpy()).max()) <= 1e-7


def test_library_table_equals_torchaudio():
    """hft_resample_build_table (host-only entry point of libhft_sm100.so) against torchaudio's own table and the oracle's."""
    import ctypes
    import math
    from nylon_amt_b200 import _lib
    from torchaudio.functional.functional import _get_sinc_resample_kernel
    L = _lib.lib()
    for sr in (44100, 48000, 22050, 8000, 32000, 11025):
        o, n, w = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32()
        need = L.hft_resample_build_table(sr, 16000, None, 0, ctypes.byref(o), ctypes.byref(n), ctypes.byref(w))
        buf = np.zeros(need, np.float32)
        assert L.hft_resample_build_table(sr, 16000, buf.ctypes.data_as(ctypes.c_void_p), need, ctypes.byref(o), ctypes.byref(n), ctypes.byref(w)) == need
        g = math.gcd(sr, 16000)
        k, width = _get_sinc_resample_kernel(sr, 16000, g)
        assert (o.value, n.value, w.value) == (sr // g, 16000 // g, width)
        mine = buf.reshape(n.value, 2 * w.value + o.value)
        assert float(np.abs(mine - k.reshape(n.value, -1).numpy()).max()) <= 1e-7, sr
        assert float(np.abs(mine - resample_oracle.sinc_kernel(sr, 16000)[0]).max()) <= 1e-7, sr


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_cuda_resample_matches_reference(fx, name):
    import nylon_amt_b200 as hft
    amt = hft.AMT(hft.default_config(), None, None)
    wave = torch.from_numpy(fx["pcm_" + name].astype(np.float32) / np.float32(32768.0)).cuda()
    y = amt.wave2mono16k(wave, int(fx["sr_" + name])).cpu().numpy()
    ref = fx["mono16k_" + name]
    assert y.shape == ref.shape
    assert float(np.abs(y - ref).max()) <= 2e-6, name          # fp32 summation order over <= 475 taps of |x| <= 0.5


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_wav2feature_on_non_16k_files(fx, name, tmp_path):
    """The whole amt.py:55-61 chain from a wav file at another rate / channel count."""
    import nylon_amt_b200 as hft
    amt = hft.AMT(hft.default_config(), None, None)
    pcm = fx["pcm_" + name]
    p = str(tmp_path / (name + ".wav"))
    with _wave.open(p, "wb") as f:
        f.setnchannels(pcm.shape[0]); f.setsampwidth(2); f.setframerate(int(fx["sr_" + name]))
        f.writeframes(np.ascontiguousarray(pcm.T).tobytes())
    feat = amt.wav2feature(p)
    ref = fx["feat_" + name]
    assert tuple(feat.shape) == ref.shape and feat.dtype == torch.float32 and not feat.is_cuda
    ok, worst = logmel_oracle.close_logmel(feat.numpy(), ref, 1e-4, fft_noise=256.0)      # tonal clips: same allowance as the log-mel fixtures
    assert ok, (name, worst)


@pytest.mark.gpu
def test_resample_argument_errors():
    import ctypes
    import nylon_amt_b200 as hft
    from nylon_amt_b200 import _lib
    amt = hft.AMT(hft.default_config(), None, None)
    plan = amt._resample_plan(44100)
    x = torch.zeros((1, 441), device="cuda")
    out = torch.zeros(10, device="cuda")
    rc = _lib.lib().hft_resample_mono_f32(plan.ptr, ctypes.c_void_p(x.data_ptr()), 1, 441, ctypes.c_void_p(out.data_ptr()), 10, None)
    assert rc != 0 and b"expected 160" in _lib.lib().hft_last_error()

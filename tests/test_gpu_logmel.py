"""GPU parity of the fused log-mel kernel (through the C ABI) against the reference goldens and the CPU oracle."""
import ctypes
import os

import numpy as np
import pytest
import torch

import nylon_amt_b200 as hft
from oracle import c_logmel
from oracle import logmel_oracle as lo

pytestmark = pytest.mark.gpu
TONAL = ("sines", "sines_noise")


@pytest.fixture(scope="module")
def amt():
    assert torch.cuda.is_available()
    return hft.AMT(hft.default_config(), None, None)


def _cases(golden_dir):
    g = np.load(os.path.join(golden_dir, "logmel.npz"))
    for k in g.files:
        if k.startswith("pcm_"):
            yield k[4:], g[k].astype(np.float32) / np.float32(32768.0), g["feat_" + k[4:]]


def test_matches_reference_golden(amt, golden_dir):
    """Outputs of the reference's AMT.wav2feature (amt.py:34-63).  Tolerance: |a-b| <= 1e-4*max(1,|b|) (north_star);
    tonal clips add the fp32-FFT dynamic-range allowance documented in oracle.logmel_oracle.close_logmel."""
    for name, x, ref in _cases(golden_dir):
        out = amt.wave2feature(torch.from_numpy(x).cuda()).cpu().numpy()
        assert out.shape == ref.shape, name
        ok, worst = lo.close_logmel(out, ref, tol=1e-4, fft_noise=256.0 if name in TONAL else 0.0)
        assert ok, (name, worst)
        miss, cells, strict = lo.strict_misses(out, ref)
        if name in TONAL:              # how far the allowance reaches: cells that miss the strict 1e-4 rule (all at the log(1e-8) floor under a pure tone)
            print("log-mel %-12s strict-rule misses %d / %d cells, worst strict ratio %.2e" % (name, miss, cells, strict))
        else:
            assert miss == 0


def test_wav2feature_file_api(amt, golden_dir, tmp_path):
    """The named entry point: wav path in, CPU FloatTensor [T,256] out (amt.py:34,63), host buffers through the ABI."""
    import wave
    g = np.load(os.path.join(golden_dir, "logmel.npz"))
    p = str(tmp_path / "clip.wav")
    with wave.open(p, "wb") as f:
        f.setnchannels(1); f.setsampwidth(2); f.setframerate(16000); f.writeframes(g["pcm_randn"].astype("<i2").tobytes())
    out = amt.wav2feature(p)
    assert isinstance(out, torch.Tensor) and out.device.type == "cpu" and out.dtype == torch.float32
    ok, worst = lo.close_logmel(out.numpy(), g["feat_randn"], 1e-4)
    assert ok, worst


@pytest.mark.parametrize("n", [0, 1, 255, 256, 257, 1023, 1024, 2048, 4095, 4096, 4097, 16 * 256 - 1, 16 * 256, 16 * 256 + 1, 100003])
def test_ragged_lengths_vs_oracle(amt, mel_tables, n):
    fb, win = mel_tables
    rng = np.random.default_rng(n)
    x = (0.3 * rng.standard_normal(n)).astype(np.float32)
    out = amt.wave2feature(torch.from_numpy(x).cuda()).cpu().numpy()
    ref = c_logmel.logmel(x, win, fb)
    assert out.shape == ref.shape == (1 + n // 256, 256)
    ok, worst = lo.close_logmel(out, ref, 1e-4)
    assert ok, (n, worst)


def test_unaligned_device_pointer(amt, mel_tables):
    """A waveform that is not 16-byte aligned cannot use the TMA path; the guarded path must give the same rows."""
    fb, win = mel_tables
    rng = np.random.default_rng(5)
    x = (0.3 * rng.standard_normal(50001)).astype(np.float32)
    base = torch.from_numpy(x).cuda()
    a = amt.wave2feature(base[1:]).cpu().numpy()
    b = c_logmel.logmel(x[1:], win, fb)
    ok, worst = lo.close_logmel(a, b, 1e-4)
    assert ok, worst


def test_silence_and_floor(amt):
    out = amt.wave2feature(torch.zeros(5000, device="cuda"))
    assert out.shape == (20, 256)
    assert torch.all(out == float(np.log(np.float32(1e-8))))


def test_batch_equals_single(amt):
    rng = np.random.default_rng(11)
    lens = [1000, 0, 70001, 4096, 33333, 255]
    waves = [torch.from_numpy((0.2 * rng.standard_normal(n)).astype(np.float32)).cuda() for n in lens]
    outs = amt.waves2features(waves)
    for w, o in zip(waves, outs):
        single = amt.wave2feature(w)
        assert o.shape == single.shape
        assert torch.equal(o, single)            # same kernel, same arithmetic: bit-identical


def test_full_size_properties(amt, mel_tables):
    """BASELINE config 2 size (1 h = 57.6 M samples -> 225 001 frames).  Size-independent properties: frames are
    independent, so any window of the long clip equals the same samples transformed alone (bit-exact), and a
    sampled subset matches the CPU oracle."""
    fb, win = mel_tables
    g = torch.Generator(device="cuda").manual_seed(1234)
    x = 0.1 * torch.randn(57_600_000, device="cuda", generator=g)
    out = amt.wave2feature(x)
    assert out.shape == (225001, 256) and bool(torch.isfinite(out).all())
    nfr = 300
    for t0 in (0, 1000, 112345, 224000):
        s0 = max(0, t0 * 256 - 1024)                    # support of frames t0 .. t0+nfr-1
        s1 = min(x.numel(), (t0 + nfr - 1) * 256 + 1024)
        sub = amt.wave2feature(x[s0:s1].clone())
        shift = s0 // 256                               # sub-clip frame f is frame shift+f of the long clip
        lo_f = 0 if s0 == 0 else 4                      # frames whose 2048-sample support lies inside the excerpt
        hi_f = sub.shape[0] - 4
        assert hi_f - lo_f > 200
        assert torch.equal(out[shift + lo_f: shift + hi_f], sub[lo_f:hi_f]), t0
    xs = x[:16000 * 20].cpu().numpy()
    ref = c_logmel.logmel(xs, win, fb)
    ok, worst = lo.close_logmel(out[:1000].cpu().numpy(), ref[:1000], 1e-4)
    assert ok, worst


def test_c_abi_argument_errors(amt):
    from nylon_amt_b200 import _lib
    plan = amt._logmel_plan()
    x = torch.zeros(1000, device="cuda")
    out = torch.zeros((10, 256), device="cuda")
    rc = _lib.lib().hft_logmel_f32(plan.ptr, ctypes.c_void_p(x.data_ptr()), 1000, ctypes.c_void_p(out.data_ptr()), 10, None)
    assert rc == 10001 and b"n_frames" in _lib.lib().hft_last_error()


def test_conv_wav2fe_driver(amt, golden_dir, tmp_path):
    """The reference's corpus driver (conv_wav2fe.py:13-50) with the same command line: lists -> pickled CPU feature tensors."""
    import json
    import pickle
    import wave as _wave
    from nylon_amt_b200 import conv_wav2fe
    g = np.load(os.path.join(golden_dir, "logmel.npz"))
    names = [k[4:] for k in g.files if k.startswith("pcm_")][:3]
    d_list, d_wav, d_out = tmp_path / "list", tmp_path / "wav", tmp_path / "feature"
    for d in (d_list, d_wav, d_out):
        d.mkdir()
    for n in names:
        pcm = g["pcm_" + n]
        with _wave.open(str(d_wav / (n + ".wav")), "wb") as f:
            f.setnchannels(1); f.setsampwidth(2); f.setframerate(16000)
            f.writeframes(np.ascontiguousarray(pcm).astype("<i2").tobytes())
    (d_list / "train.list").write_text("\n".join(names[:2]) + "\n")
    (d_list / "test.list").write_text(names[2] + "\n")
    (d_list / "valid.list").write_text("")
    cfg = tmp_path / "config.json"
    cfg.write_text(json.dumps(hft.default_config()))
    n = conv_wav2fe.main(["-d_list", str(d_list), "-d_wav", str(d_wav), "-d_feature", str(d_out), "-config", str(cfg)])
    assert n == 3
    for name in names:
        with open(d_out / (name + ".pkl"), "rb") as f:
            feat = pickle.load(f)
        ref = g["feat_" + name]
        assert isinstance(feat, torch.Tensor) and not feat.is_cuda and tuple(feat.shape) == ref.shape
        ok, worst = lo.close_logmel(feat.numpy(), ref, 1e-4, fft_noise=256.0)
        assert ok, (name, worst)

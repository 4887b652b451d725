"""Import the UNMODIFIED reference (/root/reference/hftt_code) as a Python package  --  build
container only.  TEST INFRASTRUCTURE.

/root/reference does not exist on the GPU box, so nothing that runs there may import this module;
it is used by oracle/make_golden.py (which writes tests/golden/) and by the `needs_reference` tests
that skip when the tree is absent.

Two import-time shims are needed, neither touches the hot-path arithmetic:
  * amt.py:7 imports pretty_midi (not installed)  -> empty stub module (only note2midi uses it)
  * torchaudio.load (amt.py:55) needs TorchCodec in torchaudio 2.11 -> a stdlib `wave` reader that
    returns (float32[C,N] = int16/32768, sr) exactly like torchaudio's default normalisation.
"""
import os
import sys
import types
import wave as _wave

import numpy as np

REF_ROOT = os.environ.get("NYLON_REF_ROOT", "/root/reference")
REF_CODE = os.path.join(REF_ROOT, "hftt_code")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_CODE, "model", "amt.py"))


def _wav_load(path):
    import torch
    with _wave.open(path, "rb") as f:
        n_ch, width, sr, n = f.getnchannels(), f.getsampwidth(), f.getframerate(), f.getnframes()
        raw = f.readframes(n)
    assert width == 2, "shim reads 16-bit PCM only"
    a = np.frombuffer(raw, dtype="<i2").reshape(-1, n_ch).T.astype(np.float32) / np.float32(32768.0)
    return torch.from_numpy(np.ascontiguousarray(a)), sr


def load():
    """Returns (amt_module, model_spec2midi_module) of the reference."""
    if not available():
        raise RuntimeError("reference tree not present at " + REF_ROOT)
    if "pretty_midi" not in sys.modules:
        try:
            import pretty_midi  # noqa: F401
        except Exception:
            sys.modules["pretty_midi"] = types.ModuleType("pretty_midi")
    import torchaudio
    torchaudio.load = _wav_load
    if REF_CODE not in sys.path:
        sys.path.insert(0, REF_CODE)
    from model import amt as ref_amt
    from model import model_spec2midi as ref_model
    return ref_amt, ref_model


def load_train():
    """The reference's training/train.py (train(), valid()) as a module; mir_eval (absent here, used only when valid(metrics=True)) is
    stubbed the way pretty_midi is."""
    if not os.path.isfile(os.path.join(REF_CODE, "training", "train.py")):
        raise RuntimeError("reference training/train.py not present under " + REF_ROOT)
    if "mir_eval" not in sys.modules:
        try:
            import mir_eval  # noqa: F401
        except Exception:
            sys.modules["mir_eval"] = types.ModuleType("mir_eval")
    import importlib.util
    spec = importlib.util.spec_from_file_location("hftt_reference_train", os.path.join(REF_CODE, "training", "train.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def config():
    import json
    with open(os.path.join(REF_CODE, "corpus", "config.json"), "r", encoding="utf-8") as f:
        cfg = json.load(f)
    # injected by make_dataset.py:274-278,305-308 in the reference pipeline
    cfg["input"]["min_value"] = float(np.log(np.float32(1e-8)))
    cfg["input"]["max_value"] = 0.0
    return cfg


def build_model(ref_model, cfg, hid_dim, pf_dim, n_layers, n_heads, seed=1234, cnn_channel=4, cnn_kernel=5):
    """Construction + init exactly as m_training.py:110-141 (xavier_uniform on every weight with dim>1)."""
    import torch
    import torch.nn as nn
    torch.manual_seed(seed)
    dev = "cpu"
    enc = ref_model.Encoder_SPEC2MIDI(cfg["input"]["margin_b"], cfg["input"]["num_frame"], cfg["feature"]["n_bins"],
                                      cnn_channel, cnn_kernel, hid_dim, n_layers, n_heads, pf_dim, 0.1, dev)
    dec = ref_model.Decoder_SPEC2MIDI(cfg["input"]["num_frame"], cfg["feature"]["n_bins"], cfg["midi"]["num_note"],
                                      cfg["midi"]["num_velocity"], hid_dim, n_layers, n_heads, pf_dim, 0.1, dev)
    model = ref_model.Model_SPEC2MIDI(enc, dec)

    def initialize_weights(m):
        if hasattr(m, "weight") and m.weight.dim() > 1:
            nn.init.xavier_uniform_(m.weight.data)
    model.apply(initialize_weights)
    model.eval()
    return model


def write_wav16(path, x, sr=16000):
    """float [-1,1) -> 16-bit PCM mono wav (the synthetic-set file format)."""
    q = np.clip(np.round(np.asarray(x, dtype=np.float64) * 32768.0), -32768, 32767).astype("<i2")
    with _wave.open(path, "wb") as f:
        f.setnchannels(1)
        f.setsampwidth(2)
        f.setframerate(sr)
        f.writeframes(q.tobytes())
    return q.astype(np.float32) / np.float32(32768.0)

"""nylon_amt_b200 -- B200 (sm_100a) implementation of nylon-amt's hFT-Transformer transcription hot path:
AMT.wav2feature (log-mel) and Model_SPEC2MIDI.forward, behind the reference's own Python signatures.

    import nylon_amt_b200 as hft
    amt = hft.AMT(config, None, None)                 # same constructor as hftt_code/model/amt.py:10
    feature = amt.wav2feature("clip.wav")             # fused CUDA log-mel kernel

The compute lives in libhft_sm100.so (C ABI: include/hft_sm100.h); see DESIGN.md and INTEGRATION.md.
"""
import sys
import types

from ._lib import lib, LIB_PATH  # noqa: F401


def default_config():
    """hftt_code/corpus/config.json:1-24 plus the two values hftt_code/corpus/make_dataset.py:274-278 injects."""
    return {
        "feature": {"sr": 16000, "hop_sample": 256, "mel_bins": 256, "n_bins": 256, "fft_bins": 2048, "window_length": 2048,
                    "log_offset": 1e-8, "window": "hann", "pad_mode": "constant"},
        "input": {"margin_b": 32, "margin_f": 32, "num_frame": 128, "min_value": -18.42068099975586,   # float(np.log(np.float32(1e-8)))
                  "max_value": 0.0},
        "midi": {"note_min": 21, "note_max": 108, "num_note": 88, "num_velocity": 128},
    }


def __getattr__(name):
    # lazy: importing the package must not import torch-heavy modules until they are used
    if name == "AMT":
        from .amt import AMT
        return AMT
    if name in ("Model_SPEC2MIDI", "Encoder_SPEC2MIDI", "Decoder_SPEC2MIDI"):
        from . import model_spec2midi
        return getattr(model_spec2midi, name)
    if name == "training":
        import importlib
        return importlib.import_module(".training", __name__)
    raise AttributeError(name)


def build_model(config, hid_dim=256, pf_dim=512, n_layers=3, n_heads=4, cnn_channel=4, cnn_kernel=5, dropout=0.1, seed=None,
                device="cuda"):
    """Construct + initialise exactly like hftt_code/training/m_training.py:117-141 (xavier_uniform on every
    weight with dim > 1, including embeddings); returns the module in eval mode on `device`."""
    import torch
    import torch.nn as nn
    from .model_spec2midi import Encoder_SPEC2MIDI, Decoder_SPEC2MIDI, Model_SPEC2MIDI
    if seed is not None:
        torch.manual_seed(seed)
    enc = Encoder_SPEC2MIDI(config["input"]["margin_b"], config["input"]["num_frame"], config["feature"]["n_bins"], cnn_channel,
                            cnn_kernel, hid_dim, n_layers, n_heads, pf_dim, dropout, device)
    dec = Decoder_SPEC2MIDI(config["input"]["num_frame"], config["feature"]["n_bins"], config["midi"]["num_note"],
                            config["midi"]["num_velocity"], hid_dim, n_layers, n_heads, pf_dim, dropout, device)
    model = Model_SPEC2MIDI(enc, dec)

    def initialize_weights(m):
        if hasattr(m, "weight") and m.weight.dim() > 1:
            nn.init.xavier_uniform_(m.weight.data)
    model.apply(initialize_weights)
    model = model.to(device)
    model.eval()
    return model


def install_reference_aliases():
    """Make `model.model_spec2midi` / `model.amt` importable names that resolve to this package, so modules pickled
    by the reference (pickle stores the class path, amt.py:24-25) load into the B200 implementation."""
    from . import amt as _amt
    from . import model_spec2midi as _msm
    if "model" not in sys.modules:
        pkg = types.ModuleType("model")
        pkg.__path__ = []
        sys.modules["model"] = pkg
    sys.modules["model"].model_spec2midi = _msm
    sys.modules["model"].amt = _amt
    sys.modules.setdefault("model.model_spec2midi", _msm)
    sys.modules.setdefault("model.amt", _amt)

"""CPU-side checks of the boundary: the C-ABI library loads and exports every symbol include/hft_sm100.h declares,
the schema the library reports equals the reference's state_dict, tables are bit-identical to torchaudio's, the
restructured mpe2note equals the reference's output, and nothing falls back to the CPU."""
import json
import os
import re

import numpy as np
import pytest
import torch

import nylon_amt_b200 as hft
from nylon_amt_b200 import _lib, melfb
from nylon_amt_b200.notes import mpe2note
from oracle import hft_oracle as ho

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "hft_sm100.h")).read()
    declared = set(re.findall(r"\b(hft_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    L = _lib.lib()                       # dlopen + getattr on every symbol
    assert L.hft_version() == 1
    assert L.hft_logmel_num_frames(480000) == 1876 and L.hft_logmel_num_frames(0) == 1


def test_c_schema_equals_reference_state_dict(golden_dir):
    g = np.load(os.path.join(golden_dir, "hft_reduced.npz"))
    ref_keys = [k[2:] for k in g.files if k.startswith("w:")]
    model = hft.build_model(hft.default_config(), 64, 128, 2, 2, seed=1234, device="cpu")
    assert list(model.state_dict().keys()) == ref_keys
    h = model._handle()                  # hft_model_create needs no GPU
    assert h.names == ref_keys
    assert h.numel == [int(g["w:" + k].size) for k in ref_keys]
    # same seed + same construction order => bit-identical weights to the reference's (m_training.py:117-141)
    for k, v in model.state_dict().items():
        assert np.array_equal(v.numpy(), g["w:" + k]), k


def test_paper_size_weights_reproduce_reference_checksums(golden_dir):
    g = np.load(os.path.join(golden_dir, "hft_paper.npz"))
    cs = json.loads(str(g["checksums"]))
    model = hft.build_model(hft.default_config(), 256, 512, 3, 4, seed=1234, device="cpu")
    sd = model.state_dict()
    assert list(sd.keys()) == list(cs.keys()) and len(sd) == 165
    assert sum(v.numel() for v in sd.values()) == 5516574
    for k, v in sd.items():
        assert abs(float(v.double().sum()) - cs[k][0]) <= 1e-9 * max(1.0, cs[k][1]), k


def test_unsupported_dims_fail_loudly():
    cfg = hft.default_config()
    model = hft.build_model(cfg, 96, 128, 1, 2, seed=0, device="cpu")       # head_dim 48
    with pytest.raises(RuntimeError, match="head_dim"):
        model._handle()


def test_no_cpu_fallback():
    model = hft.build_model(hft.default_config(), 64, 128, 2, 2, seed=1, device="cpu")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        model(torch.zeros(1, 256, 192))
    model.train()                      # train mode runs the CUDA training forward (hft_train_forward): no CPU fallback either
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        model(torch.zeros(1, 256, 192))


def test_tables_bit_identical_to_torchaudio(mel_tables):
    fb, win = mel_tables
    assert np.array_equal(melfb.melscale_fbanks().numpy(), fb)
    assert np.array_equal(melfb.hann_window().numpy(), win)


def test_mpe2note_equals_reference(golden_dir):
    cfg = hft.default_config()
    d = np.load(os.path.join(golden_dir, "transcript_reduced.npz"))
    cases = [("notes_A", {}, "A"), ("notes_B", {}, "B"), ("notes_B_offset", dict(mode_offset="offset"), "B"),
             ("notes_B_longer", dict(thred_onset=0.4, thred_offset=0.6, thred_mpe=0.45, mode_velocity="org", mode_offset="longer"), "B")]
    for key, kw, p in cases:
        ref = json.loads(str(d[key]))
        mine = mpe2note(cfg, d["t_onset_" + p], d["t_offset_" + p], d["t_mpe_" + p], d["t_velocity_" + p], **kw)
        assert mine == ref, key


def test_mpe2note_hand_cases():
    cfg = hft.default_config()
    T = 12
    on = np.zeros((T, 88), np.float32); off = np.zeros((T, 88), np.float32); mpe = np.zeros((T, 88), np.float32)
    vel = np.full((T, 88), 64, np.int8)
    on[2, 0], on[3, 0], on[4, 0] = 0.6, 0.9, 0.7          # peak at 3, skewed right
    mpe[2:8, 0] = 0.9
    off[9, 0] = 0.8
    on[5, 1] = on[6, 1] = 0.8                             # plateau: both frames are peaks
    notes = mpe2note(cfg, on, off, mpe, vel)
    p21 = [n for n in notes if n["pitch"] == 21]
    assert len(p21) == 1 and abs(p21[0]["onset"] - (3 * 0.016 + 0.016 * 0.5 * 0.1 / 0.3)) < 1e-6
    assert p21[0]["offset"] == 8 * 0.016                  # mpe drops at frame 8 before the offset peak at 9
    assert len([n for n in notes if n["pitch"] == 22]) == 2
    assert mpe2note(cfg, on * 0, off, mpe, vel) == []


def test_window_view_matches_reference_padding(golden_dir):
    """AMT.transcript builds its windows as a strided view of the padded feature; check it against the oracle's
    restatement of amt.py:70-73,88-89 (CPU tensors, no kernel)."""
    d = np.load(os.path.join(golden_dir, "transcript_reduced.npz"))
    feat = torch.from_numpy(d["feature"])
    T, F, mb, mf = feat.shape[0], 128, 32, 32
    len_s = int(np.ceil(T / F) * F) - T
    a_input = torch.full((mb + T + len_s + mf, 256), hft.default_config()["input"]["min_value"])
    a_input[mb:mb + T] = feat
    view = torch.as_strided(a_input, (3, 256, 192), (128 * 256, 1, 256))
    assert torch.equal(view, ho.segment_feature(d["feature"]))

"""GPU: device-assisted note decoding (hft_note_* + notes.mpe2note_device) must return exactly the note list of the host restructuring
(notes.mpe2note), which tests/test_host_logic.py pins to the reference's AMT.mpe2note on the golden transcripts."""
import json
import os
import time

import numpy as np
import pytest

import nylon_amt_b200 as hft
from nylon_amt_b200 import notes

pytestmark = pytest.mark.gpu
CFG = hft.default_config()


def _random_maps(T, seed, plateau=True):
    """Activation maps with every edge case of amt.py:196-222: plateaus (equal neighbours, saturated 1.0 runs), peaks at the clip edges,
    ties between neighbours, overlapping notes, zero velocities."""
    rng = np.random.default_rng(seed)
    N = 88
    on = (rng.random((T, N)) ** 6).astype(np.float32)
    off = (rng.random((T, N)) ** 6).astype(np.float32)
    mpe = (rng.random((T, N)) ** 0.5).astype(np.float32)
    if plateau:
        q = lambda x: np.round(x * 8).astype(np.float32) / np.float32(8)          # heavy quantisation: many equal neighbours and plateaus
        on[:, ::3] = q(on[:, ::3]); off[:, 1::3] = q(off[:, 1::3])
        on[5:9, 7] = 1.0; on[0, 11] = 0.9; on[T - 1, 12] = 0.95; off[T - 1, 11] = 0.8
    vel = rng.integers(0, 128, (T, N)).astype(np.int8)
    vel[rng.random((T, N)) < 0.2] = 0
    return on, off, mpe, vel


@pytest.mark.parametrize("T,seed", [(1, 0), (2, 1), (3, 2), (50, 3), (777, 4), (4096, 5)])
@pytest.mark.parametrize("mode_offset", ["shorter", "longer", "offset"])
def test_device_equals_host_on_edge_cases(T, seed, mode_offset):
    on, off, mpe, vel = _random_maps(T, seed)
    for mode_velocity in ("ignore_zero", "org"):
        for thr in (0.5, 0.3):
            ref = notes.mpe2note(CFG, on, off, mpe, vel, thr, thr, 0.5, mode_velocity, mode_offset)
            got = notes.mpe2note_device(CFG, on, off, mpe, vel, thr, thr, 0.5, mode_velocity, mode_offset)
            assert got == ref, (T, seed, mode_offset, mode_velocity, thr, len(got), len(ref))


def test_device_equals_reference_golden(golden_dir):
    t = np.load(os.path.join(golden_dir, "transcript_reduced.npz"))
    out = [t["t_" + n] for n in ("onset_B", "offset_B", "mpe_B", "velocity_B")]
    got = notes.mpe2note_device(CFG, *out)
    ref = json.loads(str(t["notes_B"]))
    assert got == ref
    # the other parameterisations the reference wrote into the fixture (oracle/make_golden.py)
    assert notes.mpe2note_device(CFG, *out, thred_onset=0.4, thred_offset=0.6, thred_mpe=0.45, mode_velocity="org", mode_offset="longer") == \
        json.loads(str(t["notes_B_longer"]))
    assert notes.mpe2note_device(CFG, *out, mode_offset="offset") == json.loads(str(t["notes_B_offset"]))
    outA = [t["t_" + n] for n in ("onset_A", "offset_A", "mpe_A", "velocity_A")]
    assert notes.mpe2note_device(CFG, *outA) == json.loads(str(t["notes_A"]))


def test_full_hour_speed_and_identity():
    """One hour of frames (225 001 x 88) with ~18 000 notes: identical lists, and the device path is the one AMT.mpe2note picks."""
    T = 225001
    rng = np.random.default_rng(0)
    on = np.zeros((T, 88), np.float32); off = np.zeros((T, 88), np.float32); mpe = np.zeros((T, 88), np.float32); vel = np.zeros((T, 88), np.int8)
    for _ in range(18000):
        f = int(rng.integers(5, T - 200)); p = int(rng.integers(0, 88)); d = int(rng.integers(5, 150))
        on[f - 1:f + 2, p] = [0.4, 0.9, 0.5]; off[f + d - 1:f + d + 2, p] = [0.3, 0.8, 0.4]; mpe[f:f + d, p] = 0.9; vel[f, p] = rng.integers(1, 127)
    amt = hft.AMT(CFG, None, None)
    notes.mpe2note_device(CFG, on[:1000], off[:1000], mpe[:1000], vel[:1000])        # warm-up (library load, allocator)
    t0 = time.perf_counter(); got = amt.mpe2note(a_onset=on, a_offset=off, a_mpe=mpe, a_velocity=vel); t_dev = time.perf_counter() - t0
    t0 = time.perf_counter(); ref = notes.mpe2note(CFG, on, off, mpe, vel); t_host = time.perf_counter() - t0
    assert got == ref and len(ref) > 15000
    print("mpe2note 1 h: device-assisted %.2f s, host %.2f s" % (t_dev, t_host))
    assert t_dev < t_host


def test_m_inference_driver_end_to_end(golden_dir, tmp_path):
    """The evaluation driver (reference hftt_code/evaluation/m_inference.py, same flags and files): pickled model -> wav -> feature ->
    activation maps -> note json, all through the B200 path; the notes equal AMT.mpe2note on the maps it wrote."""
    import pickle
    import wave as _wave
    import torch
    from nylon_amt_b200 import m_inference
    g = np.load(os.path.join(golden_dir, "hft_reduced.npz"))
    t = np.load(os.path.join(golden_dir, "transcript_reduced.npz"))
    model = hft.build_model(CFG, 64, 128, 2, 2, device="cpu")
    model.load_state_dict({k[2:]: torch.from_numpy(g[k]).clone() for k in g.files if k.startswith("w:")})
    dirs = {k: tmp_path / k for k in ("cp", "wav", "fe", "mpe", "note")}
    for d in dirs.values():
        d.mkdir()
    with open(dirs["cp"] / "best_model.pkl", "wb") as f:
        pickle.dump(model, f)
    pcm = t["pcm"].astype("<i2")
    with _wave.open(str(dirs["wav"] / "clip.wav"), "wb") as f:
        f.setnchannels(1); f.setsampwidth(2); f.setframerate(16000); f.writeframes(pcm.tobytes())
    (tmp_path / "test.list").write_text("clip\n")
    (tmp_path / "config.json").write_text(json.dumps(CFG))
    n = m_inference.main(["-f_config", str(tmp_path / "config.json"), "-f_list", str(tmp_path / "test.list"), "-d_cp", str(dirs["cp"]), "-d_wav", str(dirs["wav"]),
                          "-d_fe", str(dirs["fe"]), "-d_mpe", str(dirs["mpe"]), "-d_note", str(dirs["note"]), "-calc_feature", "-calc_transcript"])
    assert n == 1
    with open(dirs["mpe"] / "clip_2nd.onset", "rb") as f:
        on = pickle.load(f)
    assert on.dtype == np.float32 and on.shape[1] == 88
    maps = []
    for head in ("onset", "offset", "mpe", "velocity"):
        with open(dirs["mpe"] / ("clip_2nd." + head), "rb") as f:
            maps.append(pickle.load(f))
    written = json.loads((dirs["note"] / "clip_2nd.json").read_text())
    assert written == notes.mpe2note(CFG, *maps)
    assert (dirs["note"] / "clip_1st.json").is_file() and (dirs["fe"] / "clip.pkl").is_file()

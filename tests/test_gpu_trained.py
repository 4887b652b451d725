"""GPU parity on TRAINED-LIKE paper-size weights (tests/golden/trained_paper_delta.npz: ~1 200 steps of the library's own training step
on the synthetic piano set, tools/make_trained_fixture.py) against what the UNMODIFIED reference computed with the same weights on a
held-out 20 s clip (tests/golden/trained_paper.npz, oracle/make_golden_trained.py): probabilities of AMT.transcript in every precision
mode, and the thresholded note lists of AMT.mpe2note -- pitch, onset, OFFSET and velocity -- end to end from the PCM samples."""
import json
import os

import numpy as np
import pytest
import torch

import nylon_amt_b200 as hft
import synthset

pytestmark = pytest.mark.gpu
NAMES = ["onset_A", "offset_A", "mpe_A", "velocity_A", "onset_B", "offset_B", "mpe_B", "velocity_B"]


@pytest.fixture(scope="module")
def fx(golden_dir):
    return np.load(os.path.join(golden_dir, "trained_paper.npz")), np.load(os.path.join(golden_dir, "trained_paper_delta.npz"))


def _amt(fx, decisive):
    t, delta = fx
    cfg = hft.default_config()
    model = hft.build_model(cfg, 256, 512, 3, 4, seed=1234, device="cpu")
    sd = synthset.trained_state_dict(model.state_dict(), delta)
    if decisive:
        sd = synthset.decisive_state_dict(sd, float(t["gain"]), json.loads(str(t["calib"])))
    model.load_state_dict(sd)
    amt = hft.AMT(cfg, None, batch_size=48)
    amt.model = model.cuda().eval()
    return amt


@pytest.mark.parametrize("precision,budget", [("fp32", 2e-3), ("fp16x3", 2e-3), ("mixed", 2e-2), ("fp16", 2e-2), ("bf16", 2e-2)])
def test_trained_weights_transcript_probabilities(fx, precision, budget):
    """AMT.transcript on the reference's own feature matrix: the six probability arrays within the mode's budget, velocity classes equal."""
    t, _ = fx
    amt = _amt(fx, decisive=False)
    amt.model.precision = precision
    out = amt.transcript(t["feature"])
    worst = {}
    for n, a in zip(NAMES, out):
        assert a.shape == t[n].shape and a.dtype == t[n].dtype, n
        worst[n] = float(np.abs(a.astype(np.float64) - t[n].astype(np.float64)).max()) if not n.startswith("velocity") else float((a != t[n]).mean())
    print(precision, "trained-like weights, worst abs error:", {k: "%.2e" % v for k, v in worst.items()})
    assert max(v for k, v in worst.items() if not k.startswith("velocity")) <= budget, worst
    assert max(worst["velocity_A"], worst["velocity_B"]) <= 1e-3, worst


@pytest.mark.parametrize("precision", ["fp32", "fp16x3"])
def test_trained_weights_note_lists_identical_with_offsets(fx, precision):
    """north_star: identical thresholded note lists on the synthetic set.  PCM -> AMT.wave2feature (fused log-mel) -> AMT.transcript ->
    AMT.mpe2note with the decisive re-calibration of the fixture, in the three offset modes, both heads: every note of the reference's
    lists with the same pitch, onset (sub-frame interpolated, within 1 ms), offset (within 1 ms) and velocity."""
    t, _ = fx
    amt = _amt(fx, decisive=True)
    amt.model.precision = precision
    wav = torch.from_numpy(t["pcm"].astype(np.float32) / np.float32(32768.0)).cuda()
    feat = amt.wave2feature(wav)
    assert tuple(feat.shape) == tuple(t["feature"].shape)
    out = amt.transcript(feat)
    for n, a in zip(NAMES, out):
        if not n.startswith("velocity"):
            # the thresholded maps are identical (the fixture's thresholds sit in gaps of the reference's output distribution)
            assert np.array_equal(a >= 0.5, t["g_" + n] >= 0.5), (n, int(((a >= 0.5) != (t["g_" + n] >= 0.5)).sum()))
    for key, kw, h in (("notes_A", {}, 0), ("notes_B", {}, 4), ("notes_B_offset", dict(mode_offset="offset"), 4), ("notes_B_longer", dict(mode_offset="longer"), 4)):
        ref = json.loads(str(t[key]))
        mine = amt.mpe2note(a_onset=out[h], a_offset=out[h + 1], a_mpe=out[h + 2], a_velocity=out[h + 3], mode_velocity="org", **kw)
        assert len(ref) > 100
        assert len(mine) == len(ref), (key, len(mine), len(ref))
        # the reference orders by onset time, ties by pitch (amt.py:343); two onsets 1e-5 s apart may swap, so pair the lists by (pitch, onset)
        order = lambda n: (n["pitch"], n["onset"])
        for a, b in zip(sorted(mine, key=order), sorted(ref, key=order)):
            assert a["pitch"] == b["pitch"] and a["velocity"] == b["velocity"], (key, a, b)
            assert abs(a["onset"] - b["onset"]) <= 1e-3 and abs(a["offset"] - b["offset"]) <= 1e-3, (key, a, b)

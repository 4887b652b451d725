"""The synthetic piano set -- TEST INFRASTRUCTURE (numpy only; no reference code, no product code).

north_star asks for "identical thresholded note lists on the synthetic set": this module IS that set.  A clip is a seeded random
score (pitch 21..108, onset, offset, velocity) rendered as decaying harmonic tones at 16 kHz, plus the frame labels the
reference's corpus step would derive from the same score (hftt_code/corpus/conv_note2label.py:7-110: triangular onset / offset
targets of +-3 frames, mpe while the key is down, velocity where the onset target is >= 0.5).  Used by
  * tools/make_trained_fixture.py   trains the paper-size model a few hundred steps on it (GPU box) -> tests/golden/trained_paper_delta.npz
  * oracle/make_golden_trained.py   runs the UNMODIFIED reference on a held-out clip with those weights -> tests/golden/trained_paper.npz
  * tests/test_gpu_trained.py       the CUDA path against that fixture (probabilities, note lists with offsets)
"""
import numpy as np

SR = 16000
HOP = 256
NOTE_MIN = 21
N_NOTE = 88


def random_notes(seconds, seed, notes_per_second=4.0, max_poly=6):
    """A seeded random score: list of dicts (pitch, onset [s], offset [s], velocity), sorted by onset; the same key is never struck
    while it is still down and at most `max_poly` keys sound together."""
    rng = np.random.default_rng(seed)
    n = int(seconds * notes_per_second)
    onsets = np.sort(rng.uniform(0.05, max(0.06, seconds - 0.4), n))
    notes, busy = [], {}
    for t in onsets:
        pitch = int(np.clip(np.round(rng.normal(64, 14)), NOTE_MIN, NOTE_MIN + N_NOTE - 1))
        dur = float(np.clip(rng.lognormal(-1.0, 0.6), 0.12, 1.5))
        vel = int(rng.integers(30, 115))
        if busy.get(pitch, -1.0) > t - 0.08:
            continue
        if sum(1 for e in busy.values() if e > t) >= max_poly:
            continue
        off = min(float(t + dur), seconds - 0.05)
        if off - t < 0.1:
            continue
        busy[pitch] = off
        notes.append({"pitch": pitch, "onset": float(t), "offset": off, "velocity": vel})
    return notes


def render(notes, n_samples, noise=1e-3, seed=0):
    """Decaying harmonic tones: partial k at k * f0 (inharmonicity ignored), amplitude ~ velocity / k, 4 ms attack, exponential decay
    faster for higher partials and pitches, 40 ms release at key-up; a little white noise on top.  float32 in (-1, 1)."""
    x = np.zeros(n_samples, np.float64)
    for nt in notes:
        f0 = 440.0 * 2.0 ** ((nt["pitch"] - 69) / 12.0)
        s0 = int(nt["onset"] * SR)
        s1 = min(n_samples, int((nt["offset"] + 0.2) * SR))
        if s1 <= s0:
            continue
        t = np.arange(s1 - s0) / SR
        key_up = nt["offset"] - nt["onset"]
        env = np.minimum(1.0, t / 0.004) * np.where(t < key_up, 1.0, np.exp(-(t - key_up) / 0.04))
        amp = 0.25 * (nt["velocity"] / 127.0) ** 1.5
        tone = np.zeros_like(t)
        for k in range(1, 9):
            fk = f0 * k
            if fk > 7600.0:
                break
            tone += np.sin(2 * np.pi * fk * t + 0.3 * k) * np.exp(-t * (1.2 + 0.35 * k + f0 / 900.0)) / k
        x[s0:s1] += amp * env * tone
    x += noise * np.random.default_rng(seed + 77).standard_normal(n_samples)
    peak = float(np.abs(x).max())
    if peak > 0.95:
        x *= 0.95 / peak
    return x.astype(np.float32)


def labels(notes, n_frames, tolerance=3):
    """conv_note2label.py:7-110 for hop 256 / 16 kHz (onset_tolerance = offset_tolerance = int(50 / 16 + 0.5) = 3 frames), without the
    duration-dependent offset tolerance: float32 onset / offset [n_frames, 88] in [0, 1], mpe {0, 1}, int64 velocity."""
    hop_ms = 1000.0 * HOP / SR
    fps = SR / HOP
    onset = np.zeros((n_frames, N_NOTE), np.float32)
    offset = np.zeros((n_frames, N_NOTE), np.float32)
    mpe = np.zeros((n_frames, N_NOTE), np.float32)
    velocity = np.zeros((n_frames, N_NOTE), np.int64)
    for nt in notes:
        p = nt["pitch"] - NOTE_MIN
        f_on = int(nt["onset"] * fps + 0.5)
        f_off = int(nt["offset"] * fps + 0.5)
        for j in range(-tolerance, tolerance + 1):
            f = f_on + j
            if 0 <= f < n_frames:
                v = max(0.0, 1.0 - abs(f * hop_ms - nt["onset"] * 1000.0) / (tolerance * hop_ms))
                onset[f, p] = max(onset[f, p], v)
                if onset[f, p] >= 0.5 and (j >= 0 or velocity[f, p] == 0):
                    velocity[f, p] = nt["velocity"]
            f = f_off + j
            if 0 <= f < n_frames:
                v = max(0.0, 1.0 - abs(f * hop_ms - nt["offset"] * 1000.0) / (tolerance * hop_ms))
                offset[f, p] = max(offset[f, p], v)
        mpe[max(0, f_on):min(n_frames, f_off + 1), p] = 1.0
    return onset, offset, mpe, velocity


def clip(seconds, seed, notes_per_second=4.0):
    """(waveform float32 [seconds * 16000], score, n_frames = 1 + n_samples // 256)."""
    n = int(seconds * SR)
    notes = random_notes(seconds, seed, notes_per_second)
    return render(notes, n, seed=seed), notes, 1 + n // HOP


FAMILIES = ("noise", "tonal", "silence", "fullscale", "mixed", "piano")


def family_signal(family, n=SR * 12, seed=0):
    """The signal families of the precision sweeps (tonal, silence / floor, full-scale, noise, mixed) + the piano set."""
    g = np.random.default_rng(seed)
    t = np.arange(n) / float(SR)
    if family == "noise":
        return (0.1 * g.standard_normal(n)).astype(np.float32)
    if family == "tonal":
        return (0.2 * (np.sin(2 * np.pi * 220 * t) + np.sin(2 * np.pi * 440 * t) + np.sin(2 * np.pi * 1318.5 * t))).astype(np.float32)
    if family == "silence":
        return np.concatenate([np.zeros(n // 2), 1e-3 * g.standard_normal(n - n // 2)]).astype(np.float32)
    if family == "fullscale":
        return g.uniform(-1, 1, n).astype(np.float32)
    if family == "mixed":
        x = 0.2 * np.sin(2 * np.pi * 523.25 * t) * (np.sin(2 * np.pi * 1.5 * t) > 0) + 0.003 * g.standard_normal(n)
        return x.astype(np.float32)
    if family == "piano":
        return render(random_notes(n / float(SR), 4000 + seed), n, seed=seed)
    raise ValueError(family)


def trained_state_dict(init_sd, delta):
    """The trained-like paper-size weights: init(seed 1234) + scale * int8 delta per tensor, in fp32 (tools/make_trained_fixture.py wrote
    the delta from a few hundred steps of the library's own training step on this set).  init_sd: name -> torch tensor; delta: the npz."""
    import torch
    out = {}
    for k, v in init_sd.items():
        d = delta["d:" + k].astype(np.float32) * np.float32(delta["s:" + k])
        out[k] = torch.from_numpy(v.detach().cpu().numpy().astype(np.float32) + d.reshape(tuple(v.shape)))
    return out


def decisive_state_dict(sd, gain, calib):
    """Re-calibrated, decisive sigmoid heads: logit' = gain * (logit - c_head) for the six fc_{onset,offset,mpe}_{freq,time} heads
    (weight rows x gain, bias -> gain * (bias - c_head)); calib: {"onset_A": c, ..., "mpe_B": c} (A = freq head, B = time head)."""
    import torch
    out = {k: v.clone() for k, v in sd.items()}
    for n in ("onset", "offset", "mpe"):
        for s, h in (("freq", "A"), ("time", "B")):
            w, b = "decoder_spec2midi.fc_%s_%s.weight" % (n, s), "decoder_spec2midi.fc_%s_%s.bias" % (n, s)
            out[w] = out[w] * torch.tensor(gain, dtype=out[w].dtype)
            out[b] = (out[b] - torch.tensor(calib["%s_%s" % (n, h)], dtype=out[b].dtype)) * torch.tensor(gain, dtype=out[b].dtype)
    return out

"""Generate tests/golden/*.npz from the UNMODIFIED reference  --  build container only.

Run:  python oracle/make_golden.py          (needs /root/reference; the GPU box never runs this)

Every vector below is an output of the reference's own code on a seeded input:
  * logmel_*.npz     AMT.wav2feature (hftt_code/model/amt.py:34-63) on 16-bit 16 kHz mono wavs
  * mel_fb.npz       the torchaudio filterbank / window the reference's MelSpectrogram builds (amt.py:59)
  * hft_reduced.npz  Model_SPEC2MIDI.forward (model_spec2midi.py:15-35), README reduced size, seed 1234
  * hft_paper.npz    same, paper size (weights regenerated from the seed; checksums stored)
  * transcript_reduced.npz  AMT.transcript / transcript_stride / mpe2note (amt.py:66-344) on a 6 s clip,
                     "decisive" weights
TEST INFRASTRUCTURE.
"""
import json
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import _refload  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
SIZES = {"reduced": (64, 128, 2, 2), "paper": (256, 512, 3, 4)}


def signals(n, seed):
    """The five fixture signals of SURVEY.md 8c."""
    rng = np.random.default_rng(seed)
    t = np.arange(n) / 16000.0
    s3 = 0.2 * (np.sin(2 * np.pi * 220 * t) + np.sin(2 * np.pi * 440 * t) + np.sin(2 * np.pi * 1318.5 * t))
    half = np.zeros(n)
    half[n // 2:] = 1e-3 * rng.standard_normal(n - n // 2)
    return {
        "randn": 0.1 * rng.standard_normal(n),
        "unif": rng.uniform(-1, 1, n),
        "sines": s3,
        "sines_noise": s3 + 2e-4 * rng.standard_normal(n),
        "halfsilence": half,
    }


def clip_feature(A, x, tmp):
    p = os.path.join(tmp, "clip.wav")
    xq = _refload.write_wav16(p, x)
    pcm = np.round(xq * 32768.0).astype(np.int16)
    return pcm, A.wav2feature(p).numpy()


def weight_checksums(sd):
    return {k: [float(v.double().sum()), float(v.double().abs().sum())] for k, v in sd.items()}


def sub_vel(v):
    return v[:, ::8, ::8, :].contiguous().numpy()


def sub_attn(a):
    return a[:, ::16, :, ::11, :].contiguous().numpy()


def decisive(sd, gain=8.0):
    """scale the six fc_{onset,offset,mpe}_* weight rows so head outputs are bimodal."""
    sd = {k: v.clone() for k, v in sd.items()}
    for n in ("onset", "offset", "mpe"):
        for s in ("freq", "time"):
            sd["decoder_spec2midi.fc_%s_%s.weight" % (n, s)] *= gain
    return sd


def main():
    os.makedirs(OUT, exist_ok=True)
    ref_amt, ref_model = _refload.load()
    cfg = _refload.config()
    A = ref_amt.AMT(cfg, None, None)
    tmp = tempfile.mkdtemp()

    # ---- filterbank / window the reference builds through torchaudio (amt.py:59) ----------------
    import torchaudio
    tr = torchaudio.transforms.MelSpectrogram(sample_rate=16000, n_fft=2048, win_length=2048, hop_length=256,
                                              pad_mode="constant", n_mels=256, norm="slaney")
    fb = tr.mel_scale.fb.numpy()
    start = np.zeros(256, np.int32)
    length = np.zeros(256, np.int32)
    w = []
    for m in range(256):
        nz = np.nonzero(fb[:, m])[0]
        start[m], length[m] = nz[0], nz[-1] - nz[0] + 1
        w.append(fb[nz[0]:nz[-1] + 1, m])
    np.savez_compressed(os.path.join(OUT, "mel_fb.npz"), start=start, length=length, weights=np.concatenate(w),
                        window=tr.spectrogram.window.numpy(), torchaudio_version=torchaudio.__version__,
                        torch_version=torch.__version__)

    # ---- log-mel ---------------------------------------------------------------------------------
    out = {}
    for name, x in signals(16000, 7).items():
        pcm, f = clip_feature(A, x, tmp)
        out["pcm_" + name], out["feat_" + name] = pcm, f
    for n in (1, 100, 255, 256, 257, 2047, 2048, 2049, 2559, 2560, 2561, 4096):
        x = signals(n, 100 + n)["randn"]
        pcm, f = clip_feature(A, x, tmp)
        out["pcm_len%d" % n], out["feat_len%d" % n] = pcm, f
    np.savez_compressed(os.path.join(OUT, "logmel.npz"), **out)

    # ---- config-1 clip: 30 s of 0.1*randn (torch seed 0), SURVEY.md 8d ----------------------------
    torch.manual_seed(0)
    x30 = (0.1 * torch.randn(480000)).numpy()
    pcm30, feat30 = clip_feature(A, x30, tmp)
    assert feat30.shape == (1876, 256)

    # ---- model forward ---------------------------------------------------------------------------
    from oracle.hft_oracle import segment_feature
    spec_all = segment_feature(feat30)                       # [15,256,192]
    for size, (hid, pf, L, h) in SIZES.items():
        model = _refload.build_model(ref_model, cfg, hid, pf, L, h, seed=1234)
        sd = model.state_dict()
        nb = 2 if size == "reduced" else 1
        spec = spec_all[3:3 + nb].contiguous()
        with torch.no_grad():
            o = model(spec)
            # the call shape of amt.py:89 — a non-contiguous .T view, batch 1
            o_nc = model(spec_all[5].t().contiguous().t().unsqueeze(0))
        d = {"spec": spec.numpy(), "n_heads": h, "checksums": json.dumps(weight_checksums(sd)),
             "onset_A": o[0].numpy(), "offset_A": o[1].numpy(), "mpe_A": o[2].numpy(),
             "velocity_A_sub": sub_vel(o[3]), "velocity_A_argmax": o[3].argmax(3).numpy().astype(np.int16),
             "attention_sub": sub_attn(o[4]),
             "onset_B": o[5].numpy(), "offset_B": o[6].numpy(), "mpe_B": o[7].numpy(),
             "velocity_B_sub": sub_vel(o[8]), "velocity_B_argmax": o[8].argmax(3).numpy().astype(np.int16),
             "seg5_onset_B": o_nc[5].numpy()}
        if size == "reduced":
            for k, v in sd.items():
                d["w:" + k] = v.numpy()
        np.savez_compressed(os.path.join(OUT, "hft_%s.npz" % size), **d)

    # ---- transcript / mpe2note on the reduced model with decisive weights ------------------------
    hid, pf, L, h = SIZES["reduced"]
    model = _refload.build_model(ref_model, cfg, hid, pf, L, h, seed=1234)
    model.load_state_dict(decisive(model.state_dict()))
    A2 = ref_amt.AMT(cfg, None, None)
    A2.model = model
    A2.device = "cpu"
    feat6 = feat30[:375]                                     # 6 s -> 3 segments (last one ragged)
    names = ["onset_A", "offset_A", "mpe_A", "velocity_A", "onset_B", "offset_B", "mpe_B", "velocity_B"]
    t_out = A2.transcript(feat6)
    s_out = A2.transcript_stride(feat6, 32)
    d = {"pcm": pcm30[:96000 + 0], "feature": feat6, "gain": 8.0}
    for n, a in zip(names, t_out):
        d["t_" + n] = a
    for n, a in zip(names, s_out):
        d["s_" + n] = a
    notes_A = A2.mpe2note(a_onset=t_out[0], a_offset=t_out[1], a_mpe=t_out[2], a_velocity=t_out[3])
    notes_B = A2.mpe2note(a_onset=t_out[4], a_offset=t_out[5], a_mpe=t_out[6], a_velocity=t_out[7])
    notes_B_longer = A2.mpe2note(a_onset=t_out[4], a_offset=t_out[5], a_mpe=t_out[6], a_velocity=t_out[7],
                                 thred_onset=0.4, thred_offset=0.6, thred_mpe=0.45, mode_velocity="org", mode_offset="longer")
    notes_B_offset = A2.mpe2note(a_onset=t_out[4], a_offset=t_out[5], a_mpe=t_out[6], a_velocity=t_out[7], mode_offset="offset")
    d["notes_A"], d["notes_B"] = json.dumps(notes_A), json.dumps(notes_B)
    d["notes_B_longer"], d["notes_B_offset"] = json.dumps(notes_B_longer), json.dumps(notes_B_offset)
    guard = [int((np.abs(t_out[i] - 0.5) < 0.02).sum()) for i in (0, 1, 2, 4, 5, 6)]
    d["guard_band_counts"] = np.array(guard)
    np.savez_compressed(os.path.join(OUT, "transcript_reduced.npz"), **d)
    print("notes A/B:", len(notes_A), len(notes_B), "guard-band entries:", guard)
    for f in sorted(os.listdir(OUT)):
        print("%-28s %8.1f KB" % (f, os.path.getsize(os.path.join(OUT, f)) / 1024))


if __name__ == "__main__":
    main()

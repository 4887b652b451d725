"""tests/golden/hft_paper_families.npz: the UNMODIFIED reference's paper-size forward (seeded weights, seed 1234) on two segments of each
signal family of tests/synthset.py (noise, tonal, silence / floor, full-scale, mixed, piano) -- build container only (needs /root/reference).
The inputs are the reference's own AMT.wav2feature of 16-bit wavs of those signals, segmented as amt.py:70-89 does.
TEST INFRASTRUCTURE.   Run:  python oracle/make_golden_families.py
"""
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import _refload                      # noqa: E402
from oracle.hft_oracle import segment_feature    # noqa: E402
import synthset                                  # noqa: E402


def main():
    ref_amt, ref_model = _refload.load()
    cfg = _refload.config()
    A = ref_amt.AMT(cfg, None, None)
    tmp = tempfile.mkdtemp()
    model = _refload.build_model(ref_model, cfg, 256, 512, 3, 4, seed=1234)
    specs, fam = [], []
    for f in synthset.FAMILIES:
        p = os.path.join(tmp, f + ".wav")
        _refload.write_wav16(p, synthset.family_signal(f))
        seg = segment_feature(A.wav2feature(p).numpy())
        # silence: segment 2 straddles the zero -> noise-floor boundary, segment 1 is all floor
        for s in ((1, 2) if f == "silence" else (1, 3)):
            specs.append(seg[s])
            fam.append(f)
    spec = torch.stack(specs).contiguous()
    outs = []
    with torch.no_grad():
        for i in range(spec.shape[0]):
            outs.append(model(spec[i:i + 1]))
    o = [torch.cat([x[k] for x in outs]) for k in range(9)]
    d = {"spec": spec.numpy(), "family": np.array(fam),
         "onset_A": o[0].numpy(), "offset_A": o[1].numpy(), "mpe_A": o[2].numpy(),
         "velocity_A_sub": o[3][:, ::8, ::8, :].contiguous().numpy(), "velocity_A_argmax": o[3].argmax(3).numpy().astype(np.int16),
         "attention_sub": o[4][:, ::16, :, ::11, :].contiguous().numpy(),
         "onset_B": o[5].numpy(), "offset_B": o[6].numpy(), "mpe_B": o[7].numpy(),
         "velocity_B_sub": o[8][:, ::8, ::8, :].contiguous().numpy(), "velocity_B_argmax": o[8].argmax(3).numpy().astype(np.int16)}
    # how close the two largest velocity logits are (an argmax may legitimately flip where the gap is below the parity budget)
    for h, i in (("A", 3), ("B", 8)):
        top = o[i].topk(2, dim=3).values
        d["velocity_%s_gap" % h] = (top[..., 0] - top[..., 1]).numpy().astype(np.float32)
    path = os.path.join(ROOT, "tests", "golden", "hft_paper_families.npz")
    np.savez_compressed(path, **d)
    print("%s %.1f KB, %d segments" % (path, os.path.getsize(path) / 1024, spec.shape[0]))


if __name__ == "__main__":
    main()

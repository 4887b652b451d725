import json, os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nylon_amt_b200 as hft
cfg = hft.default_config()
t = np.load("tests/golden/transcript_reduced.npz"); g = np.load("tests/golden/hft_reduced.npz")
model = hft.build_model(cfg, 64, 128, 2, 2, device="cpu")
sd = {k[2:]: torch.from_numpy(g[k]).clone() for k in g.files if k.startswith("w:")}
for n in ("onset", "offset", "mpe"):
    for s in ("freq", "time"):
        sd["decoder_spec2midi.fc_%s_%s.weight" % (n, s)] *= float(t["gain"])
model.load_state_dict(sd)
amt = hft.AMT(cfg, None, batch_size=2); amt.model = model.cuda().eval()
for prec in ("fp32", "fp16x3", "bf16"):
    amt.model.precision = prec
    out = amt.transcript(t["feature"])
    for key, idx in (("notes_A", (0,1,2,3)), ("notes_B", (4,5,6,7))):
        if key not in t.files: continue
        notes = amt.mpe2note(a_onset=out[idx[0]], a_offset=out[idx[1]], a_mpe=out[idx[2]], a_velocity=out[idx[3]])
        ref = json.loads(str(t[key]))
        k = lambda n: (n["pitch"], round(n["onset"] / 0.016))
        a, b = {k(n) for n in notes}, {k(n) for n in ref}
        exact = sum(1 for x, y in zip(notes, ref) if x == y)
        # guard band: reference probabilities within 10*tol of a threshold
        print(prec, key, "ours", len(notes), "ref", len(ref), "symdiff", len(a ^ b), "identical dict entries", exact)
print([k for k in t.files][:30])

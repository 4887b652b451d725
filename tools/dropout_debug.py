import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nylon_amt_b200 as hft
from oracle import train_oracle
g = np.load("tests/golden/hft_reduced.npz")
sd = {k[2:]: torch.from_numpy(g[k]).clone() for k in g.files if k.startswith("w:")}
spec = torch.from_numpy(g["spec"][:1]).clone()
lab = train_oracle.synthetic_labels(1, seed=21)
for p in (0.1,):
    model = hft.build_model(hft.default_config(), 64, 128, 2, 2, dropout=p, device="cuda")
    model.load_state_dict(sd)
    opt = hft.training.Adam(model, batch_size=1, seed=5)
    loss = float(opt.forward_backward(spec.cuda(), *[x.cuda() for x in lab]).item())
    ref_loss, ref_g = train_oracle.loss_and_grads_dropout(sd, 2, spec, *lab, p=p, seed=opt.last_dropout_seed)
    _, ref64 = None, None
    o = train_oracle.DropOracle(sd, 2, p, opt.last_dropout_seed, dtype=torch.float64)
    outs = o.forward_grad(spec)
    l64 = train_oracle.loss_from_outputs(outs, *[x.double() for x in lab[:3]], lab[3]); l64.backward()
    g64 = {k: v.grad for k, v in o.sd.items()}
    print("p", p, "loss", loss, ref_loss, float(l64))
    rows = []
    for name, ref in ref_g.items():
        mx = float(g64[name].abs().max()) + 1e-30
        e_cuda = float((opt.grad_of(name).cpu().double() - g64[name]).abs().max())
        e_ref = float((ref.double() - g64[name]).abs().max())
        rows.append((e_cuda / mx, e_ref / mx, mx, name))
    rows.sort(reverse=True)
    for r in rows[:14]:
        print("cuda-vs-f64 %.2e   f32oracle-vs-f64 %.2e   max %.2e  %s" % r)

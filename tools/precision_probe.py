import sys, numpy as np, torch
sys.path.insert(0, '.')
import nylon_amt_b200 as hft
from oracle import hft_oracle
dev = torch.device('cuda:0')
cfg = hft.default_config()
amt = hft.AMT(cfg, None, None)
torch.manual_seed(0)
wave = 0.1 * torch.randn(16000 * 4)
feat = amt.wave2feature(wave.to(dev))
spec = hft_oracle.segment_feature(feat.cpu().numpy())[:1]
model = hft.build_model(cfg, 256, 512, 3, 4, seed=1234, device=dev)
sd = {k: v.cpu() for k, v in model.state_dict().items()}
o32 = hft_oracle.Oracle(sd, 4)(spec)
o64 = hft_oracle.Oracle(sd, 4, dtype=torch.float64)(spec)
names = ["onset_A", "offset_A", "mpe_A", "velocity_A", "attention", "onset_B", "offset_B", "mpe_B", "velocity_B"]
res = {}
for prec in ("fp32", "fp16x3"):
    model.precision = prec
    res[prec] = [t.cpu() for t in model(spec.to(dev))]
print("%-12s %10s %10s %10s %10s %10s" % ("output", "o32-o64", "f32-o64", "x3-o64", "x3-f32", "max|o64|"))
for i, n in enumerate(names):
    r = o64[i].double()
    print("%-12s %10.2e %10.2e %10.2e %10.2e %10.2e" % (n, float((o32[i].double() - r).abs().max()), float((res["fp32"][i].double() - r).abs().max()),
          float((res["fp16x3"][i].double() - r).abs().max()), float((res["fp16x3"][i] - res["fp32"][i]).abs().max()), float(r.abs().max())))

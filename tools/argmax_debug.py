import os, sys, numpy as np, torch
sys.path.insert(0, '.')
import nylon_amt_b200 as hft
g = np.load('tests/golden/hft_paper.npz')
model = hft.build_model(hft.default_config(), 256, 512, 3, 4, seed=1234, device='cuda')
model.precision = 'fp16x3'
spec = torch.from_numpy(g['spec']).cuda()
spec = torch.cat([spec, spec.flip(0), spec * 0.5], 0)
B = spec.shape[0]
model.max_batch = 2
full = model(spec)
full2 = model(spec)
print('run-to-run identical:', [bool(torch.equal(a, b)) for a, b in zip(full, full2)])
outs = [torch.empty_like(t) for t in full]
va = [torch.full((B, 128, 88), -7, device='cuda', dtype=torch.int8) for _ in range(2)]
model.forward_into(spec, outs, want_attention=False, velocity_argmax=va)
torch.cuda.synchronize()
for k, i in enumerate((3, 8)):
    ref = full[i].argmax(3).to(torch.int8)
    bad = (va[k] != ref).nonzero()
    print('head', 'AB'[k], 'mismatches', bad.shape[0], 'logits identical', bool(torch.equal(outs[i], full[i])))
    for b, f, n in bad.tolist()[:5]:
        row = full[i][b, f, n]
        top = row.topk(3)
        print(b, f, n, 'mine', int(va[k][b, f, n]), 'torch', int(ref[b, f, n]), 'top3', top.values.tolist(), top.indices.tolist(), 'row(mine)', float(row[int(va[k][b, f, n])]))

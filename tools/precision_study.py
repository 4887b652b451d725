"""Numerics study (CPU, oracle-side): which GEMMs tolerate bf16 operands inside the 2e-2 budget."""
import sys, time, numpy as np, torch
sys.path.insert(0, '.')
from oracle import _refload, hft_oracle as ho, logmel_oracle as lo

def bf16(t): return t.bfloat16().float()
def fp16(t): return t.half().float()
def split(t, f):  # hi + lo
    hi = f(t); return hi + f(t - hi)

ref_amt, ref_model = _refload.load(); cfg = _refload.config()
torch.manual_seed(0); x = (0.1 * torch.randn(480000)).numpy()
feat = lo.logmel(x)
size = sys.argv[1] if len(sys.argv) > 1 else 'reduced'
hid, pf, L, h = {'reduced': (64, 128, 2, 2), 'paper': (256, 512, 3, 4)}[size]
nseg = int(sys.argv[2]) if len(sys.argv) > 2 else 2
spec = ho.segment_feature(feat)[:nseg]
m = _refload.build_model(ref_model, cfg, hid, pf, L, h)
sd = m.state_dict()
ref = ho.Oracle(sd, h)(spec)
names = ['onA', 'offA', 'mpeA', 'velA', 'attn', 'onB', 'offB', 'mpeB', 'velB']

def run(label, q, st=None):
    t = time.time()
    out = ho.Oracle(sd, h, gemm_in=q, store=st)(spec)
    errs = [float((a - b).abs().max()) for a, b in zip(ref, out)]
    sa = max(errs[0:3]); sb = max(errs[5:8])
    print('%-44s sigA %.2e velA %.2e attn %.2e sigB %.2e velB %.2e  (%.0fs)' % (label, sa, errs[3], errs[4], sb, errs[8], time.time() - t), flush=True)

def policy(hi_tags, lo=bf16, hi=lambda t: split(t, bf16)):
    def q(t, tag=None):
        if tag is not None and any(tag.startswith(p) for p in hi_tags): return hi(t)
        return lo(t)
    return q

run('all bf16', policy([]))
run('all bf16 + bf16 stream', policy([]), bf16)
run('all fp16', policy([], lo=fp16))
run('front split, rest bf16', policy(['front']))
run('front+time0 split', policy(['front', 'time0']))
run('front+time* split', policy(['front', 'time']))
run('front+time*+headB split', policy(['front', 'time', 'headB']))
run('front+enc0 split', policy(['front', 'enc0']))
run('front+enc* split', policy(['front', 'enc']))
run('front+dec* split', policy(['front', 'dec']))
run('front fp32, rest fp16', policy(['front'], lo=fp16, hi=lambda t: t))
run('all bf16 split(hi+lo)', policy([''],))
run('all fp16 split', policy([''], hi=lambda t: split(t, fp16)))

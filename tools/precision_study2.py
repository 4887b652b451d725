"""Which operands need the hi+lo split (CPU emulation on the oracle)."""
import sys, time, numpy as np, torch
sys.path.insert(0, '.')
from oracle import _refload, hft_oracle as ho, logmel_oracle as lo
def fp16(t): return t.half().float()
def split(t): hi = fp16(t); return hi + fp16(t - hi)
ref_amt, ref_model = _refload.load(); cfg = _refload.config()
torch.manual_seed(0); x = (0.1 * torch.randn(480000)).numpy()
feat = lo.logmel(x)
size = sys.argv[1] if len(sys.argv) > 1 else 'reduced'
hid, pf, L, h = {'reduced': (64, 128, 2, 2), 'paper': (256, 512, 3, 4)}[size]
nseg = int(sys.argv[2]) if len(sys.argv) > 2 else 2
spec = ho.segment_feature(feat)[3:3+nseg]
m = _refload.build_model(ref_model, cfg, hid, pf, L, h); sd = m.state_dict()
ref = ho.Oracle(sd, h)(spec)
def run(label, rule):
    def q(t, tag=None):
        if tag is None: return t
        site, role = tag.split(':')
        return split(t) if rule(site, role) else fp16(t)
    t0 = time.time(); out = ho.Oracle(sd, h, gemm_in=q)(spec)
    e = [float((a - b).abs().max()) for a, b in zip(ref, out)]
    print('%-52s sigA %.1e velA %.1e attn %.1e sigB %.1e velB %.1e (%.0fs)' % (label, max(e[0:3]), e[3], e[4], max(e[5:8]), e[8], time.time()-t0), flush=True)
run('all single fp16', lambda s, r: False)
run('all split (x3/x4)', lambda s, r: True)
run('activations split only (a,q,k,p,v)', lambda s, r: r != 'w')
run('weights split only', lambda s, r: r == 'w')
run('linear split, attention single', lambda s, r: r in 'aw')
run('linear + qk split, pv single', lambda s, r: r in 'awqk')
run('linear + qk + v split, p single', lambda s, r: r in 'awqkv')
run('all split except front', lambda s, r: s != 'front')
run('all split except heads', lambda s, r: not s.startswith('head'))
run('time+heads split only', lambda s, r: s.startswith('time') or s.startswith('head'))
run('enc+dec split only', lambda s, r: s.startswith('enc') or s.startswith('dec') or s=='front')

"""Per-site sensitivity of the head outputs to single-product fp16 GEMMs (CPU emulation on the oracle).

Baseline = every GEMM operand split hi+lo (the fp16x3 mode).  For every site (front, enc0.., dec0.., headA, time0.., headB)
switch ONLY that site to single-product fp16 and report the worst error of each output group; then evaluate candidate plans
(sets of sites that stay split).  usage: python tools/precision_plan.py [nseg] [family]
"""
import sys, time, numpy as np, torch
sys.path.insert(0, '.')
from oracle import _refload, hft_oracle as ho, logmel_oracle as lo


def fp16(t): return t.half().float()
def split(t): hi = fp16(t); return hi + fp16(t - hi)


def signal(family, n=16000 * 12, seed=0):
    g = np.random.default_rng(seed)
    t = np.arange(n) / 16000.0
    if family == 'noise': return (0.1 * g.standard_normal(n)).astype(np.float32)
    if family == 'tonal': return (0.2 * (np.sin(2 * np.pi * 220 * t) + np.sin(2 * np.pi * 440 * t) + np.sin(2 * np.pi * 1318.5 * t))).astype(np.float32)
    if family == 'silence': return np.concatenate([np.zeros(n // 2), 1e-3 * g.standard_normal(n - n // 2)]).astype(np.float32)
    if family == 'fullscale': return g.uniform(-1, 1, n).astype(np.float32)
    if family == 'mixed':
        x = 0.2 * np.sin(2 * np.pi * 523.25 * t) * (np.sin(2 * np.pi * 1.5 * t) > 0) + 0.003 * g.standard_normal(n)
        return x.astype(np.float32)
    raise ValueError(family)


def main():
    nseg = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    family = sys.argv[2] if len(sys.argv) > 2 else 'noise'
    ref_amt, ref_model = _refload.load(); cfg = _refload.config()
    feat = lo.logmel(signal(family))
    spec = ho.segment_feature(feat)[1:1 + nseg]
    m = _refload.build_model(ref_model, cfg, 256, 512, 3, 4); sd = m.state_dict()
    ref = ho.Oracle(sd, 4)(spec)
    sites = ['front', 'enc0', 'enc1', 'enc2', 'dec0', 'dec1', 'dec2', 'headA', 'time0', 'time1', 'time2', 'headB']

    def run(label, rule):
        def q(t, tag=None):
            if tag is None: return t
            site, role = tag.split(':')
            return split(t) if rule(site, role) else fp16(t)
        t0 = time.time(); out = ho.Oracle(sd, 4, gemm_in=q)(spec)
        e = [float((a - b).abs().max()) for a, b in zip(ref, out)]
        print('%-58s sigA %.1e velA %.1e attn %.1e sigB %.1e velB %.1e (%.0fs)' % (label, max(e[0:3]), e[3], e[4], max(e[5:8]), e[8], time.time() - t0), flush=True)
        return e

    run('all split', lambda s, r: True)
    run('all single', lambda s, r: False)
    for s0 in sites:
        run('single only at ' + s0, lambda s, r, s0=s0: s != s0)
    for s0 in ('enc0', 'time0', 'dec1'):
        run('single only at %s linear (a,w)' % s0, lambda s, r, s0=s0: not (s == s0 and r in 'aw'))
        run('single only at %s scores (q,k)' % s0, lambda s, r, s0=s0: not (s == s0 and r in 'qk'))
        run('single only at %s pv (p,v)' % s0, lambda s, r, s0=s0: not (s == s0 and r in 'pv'))


if __name__ == '__main__' and not (len(sys.argv) > 1 and sys.argv[1] == 'plans'):
    main()


def plan_rule(name):
    """split(site, role) rules of the candidate per-GEMM plans ('mixed' mode of the library = plan A)."""
    if name == 'A':      # single: time layers >= 1, both head groups, P V outside encoder layer 0, decoder scores
        return lambda s, r: not (s in ('time1', 'time2', 'headA', 'headB') or (r in 'pv' and s != 'enc0') or (r in 'qk' and s.startswith('dec')))
    if name == 'B':      # A + encoder layer 0's P V single
        return lambda s, r: not (s in ('time1', 'time2', 'headA', 'headB') or r in 'pv' or (r in 'qk' and s.startswith('dec')))
    if name == 'C':      # A without the decoder scores
        return lambda s, r: not (s in ('time1', 'time2', 'headA', 'headB') or (r in 'pv' and s != 'enc0'))
    raise ValueError(name)


def plans():
    nseg = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    ref_amt, ref_model = _refload.load(); cfg = _refload.config()
    m = _refload.build_model(ref_model, cfg, 256, 512, 3, 4); sd = m.state_dict()
    for family in ('noise', 'tonal', 'silence', 'fullscale', 'mixed'):
        feat = lo.logmel(signal(family))
        spec = ho.segment_feature(feat)[1:1 + nseg]
        ref = ho.Oracle(sd, 4)(spec)
        for name in sys.argv[3:] or ['A']:
            rule = plan_rule(name)
            def q(t, tag=None):
                if tag is None: return t
                site, role = tag.split(':')
                return split(t) if rule(site, role) else fp16(t)
            out = ho.Oracle(sd, 4, gemm_in=q)(spec)
            e = [float((a - b).abs().max()) for a, b in zip(ref, out)]
            print('%-10s plan %s  sigA %.1e velA %.1e attn %.1e sigB %.1e velB %.1e' % (family, name, max(e[0:3]), e[3], e[4], max(e[5:8]), e[8]), flush=True)


if __name__ == '__main__' and len(sys.argv) > 1 and sys.argv[1] == 'plans':
    plans()

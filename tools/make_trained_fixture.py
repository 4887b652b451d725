"""Train the paper-size hFT for a few hundred steps on the synthetic piano set with THIS library's training step (GPU box), and write the
result as a compact delta against the seeded initialisation:  gpurun_out/trained_paper_delta.npz  (int8 delta + fp32 scale per tensor).

The fixture weights are DEFINED as  w = init(seed 1234) + scale * delta_int8  in fp32 (exactly reproducible on any host), so the
reference run that writes tests/golden/trained_paper.npz (oracle/make_golden_trained.py, build container) and the GPU tests use
bit-identical parameters.  TEST INFRASTRUCTURE (fixture generator); not on the product path.

usage (GPU box): python tools/make_trained_fixture.py [seconds_of_training=240] [lr=3e-4] [batch=8]
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import nylon_amt_b200 as hft          # noqa: E402
import synthset                       # noqa: E402


def make_segments(amt, cfg, n_clips, seconds, seed0, dev):
    specs, labs = [], [[], [], [], []]
    for c in range(n_clips):
        wav, notes, T = synthset.clip(seconds, seed0 + c)
        feat = amt.wave2feature(torch.from_numpy(wav).to(dev))
        T = feat.shape[0]
        n_seg = (T + 127) // 128
        a_input = torch.full((32 + n_seg * 128 + 32, 256), cfg["input"]["min_value"], device=dev)
        a_input[32:32 + T] = feat
        specs.append(torch.as_strided(a_input, (n_seg, 256, 192), (128 * 256, 1, 256)).contiguous())
        for k, a in enumerate(synthset.labels(notes, n_seg * 128)):
            labs[k].append(torch.from_numpy(a).view(n_seg, 128, 88))
    return torch.cat(specs), [torch.cat(x).to(dev) for x in labs]


def main():
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 240.0
    lr = float(sys.argv[2]) if len(sys.argv) > 2 else 3e-4
    B = int(sys.argv[3]) if len(sys.argv) > 3 else 8
    dev = torch.device("cuda")
    cfg = hft.default_config()
    amt = hft.AMT(cfg, None, None)
    t0 = time.time()
    spec, lab = make_segments(amt, cfg, 96, 16.384, 5000, dev)
    print("dataset: %d segments in %.1f s" % (spec.shape[0], time.time() - t0), flush=True)
    model = hft.build_model(cfg, 256, 512, 3, 4, dropout=0.1, seed=1234, device=dev)
    init = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    opt = hft.training.Adam(model, lr=lr, batch_size=B)
    g = torch.Generator().manual_seed(0)
    N = spec.shape[0]
    step, t0, hist = 0, time.time(), []
    while time.time() - t0 < budget:
        perm = torch.randperm(N, generator=g)
        for i in range(0, N - B + 1, B):
            idx = perm[i:i + B].to(dev)
            loss = hft.training.train_step(model, opt, spec[idx], lab[0][idx], lab[1][idx], lab[2][idx], lab[3][idx])
            step += 1
            if step % 50 == 0:
                hist.append(float(loss.item()))
                print("step %5d  loss %.4f  %.1f s" % (step, hist[-1], time.time() - t0), flush=True)
            if time.time() - t0 >= budget:
                break
    opt.sync_to_module()
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    out = {"steps": step, "lr": lr, "batch": B, "loss_history": np.array(hist, np.float32)}
    worst = 0.0
    for k, v in sd.items():
        d = (v - init[k]).numpy()
        scale = np.float32(max(float(np.abs(d).max()), 1e-12) / 127.0)
        q = np.clip(np.round(d / scale), -127, 127).astype(np.int8)
        out["d:" + k], out["s:" + k] = q, scale
        worst = max(worst, float(np.abs(d - q.astype(np.float32) * scale).max()))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    np.savez_compressed(os.path.join(ROOT, "gpurun_out", "trained_paper_delta.npz"), **out)
    print("steps %d, final loss %.4f, quantisation error of the delta <= %.2e, file %.1f MB" %
          (step, hist[-1] if hist else float("nan"), worst, os.path.getsize(os.path.join(ROOT, "gpurun_out", "trained_paper_delta.npz")) / 1e6), flush=True)
    # how decisive the quantised model is on a held-out clip (eval mode, fp32 path)
    model.load_state_dict({k: init[k] + torch.from_numpy(out["d:" + k].astype(np.float32) * out["s:" + k]) for k in sd})
    model.eval()
    model.precision = "fp32"
    vs, vl = make_segments(amt, cfg, 1, 20.0, 9000, dev)
    o = model(vs)
    for name, i, y in (("onset", 5, vl[0]), ("offset", 6, vl[1]), ("mpe", 7, vl[2])):
        p = o[i]
        tp = float(((p >= 0.5) & (y >= 0.5)).sum()); fp = float(((p >= 0.5) & (y < 0.5)).sum()); fn = float(((p < 0.5) & (y >= 0.5)).sum())
        print("%-6s B: tp %d fp %d fn %d  max p %.3f  cells within 0.02 of the threshold: %d" % (name, tp, fp, fn, float(p.max()), int(((p - 0.5).abs() < 0.02).sum())))


if __name__ == "__main__":
    main()

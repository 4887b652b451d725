"""CPU: the training-step restatement (oracle/train_oracle.py) against the fixture the unmodified reference wrote
(tests/golden/train_reduced.npz), and the data-parallel bucket logic over gloo (world_size 2)."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import train_oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def fx(golden_dir):
    t = np.load(os.path.join(golden_dir, "train_reduced.npz"))
    g = np.load(os.path.join(golden_dir, "hft_reduced.npz"))
    sd = {k[2:]: torch.from_numpy(g[k]).clone() for k in g.files if k.startswith("w:")}
    return t, sd


def _labels(t):
    return (torch.from_numpy(t["label_onset"]), torch.from_numpy(t["label_offset"]), torch.from_numpy(t["label_mpe"]), torch.from_numpy(t["label_velocity"]))


def test_loss_and_gradients_match_reference(fx):
    t, sd = fx
    loss, grads = train_oracle.loss_and_grads(sd, 2, torch.from_numpy(t["spec"]), *_labels(t))
    assert abs(loss - float(t["loss"])) <= 1e-5 * abs(float(t["loss"]))
    assert len(grads) == sum(1 for k in t.files if k.startswith("g:")) == 115
    for k, g in grads.items():
        ref = torch.from_numpy(t["g:" + k])
        tol = 1e-4 * float(ref.abs().max()) + 1e-7
        assert float((g - ref).abs().max()) <= tol, k


def test_adam_restatement_matches_torch_optim(fx):
    t, sd = fx
    grads = {k: torch.from_numpy(t["g:" + k]) for k in sd}
    p = {k: v.clone() for k, v in sd.items()}
    m = {k: torch.zeros_like(v) for k, v in sd.items()}
    v = {k: torch.zeros_like(v) for k, v in sd.items()}
    train_oracle.adam_step(p, grads, m, v, 1, lr=float(t["lr"]))
    for k in sd:
        assert float((p[k] - torch.from_numpy(t["p1:" + k])).abs().max()) <= 2e-7, k


def test_synthetic_labels_are_seeded_and_shaped():
    a = train_oracle.synthetic_labels(2)
    b = train_oracle.synthetic_labels(2)
    assert all(torch.equal(x, y) for x, y in zip(a, b))
    assert a[0].shape == (2, 128, 88) and a[3].dtype == torch.int64 and int(a[3].max()) < 128
    assert 0.0 <= float(a[0].min()) and float(a[0].max()) <= 1.0


WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1])
import numpy as np, torch, torch.distributed as dist
from oracle import train_oracle
from nylon_amt_b200 import shard
dist.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
rank, world = dist.get_rank(), dist.get_world_size()
t = np.load(os.path.join(sys.argv[1], "tests", "golden", "train_reduced.npz"))
g = np.load(os.path.join(sys.argv[1], "tests", "golden", "hft_reduced.npz"))
sd = {k[2:]: torch.from_numpy(g[k]).clone() for k in g.files if k.startswith("w:")}
lab = [torch.from_numpy(t[n]) for n in ("label_onset", "label_offset", "label_mpe", "label_velocity")]
spec = torch.from_numpy(t["spec"])
lo, hi = shard.block_range(spec.shape[0], rank, world)           # one segment per rank
loss, grads = train_oracle.loss_and_grads(sd, 2, spec[lo:hi], *[x[lo:hi] for x in lab])
names = sorted(grads)
flat = shard.flatten_bucket([grads[n] for n in names])            # ONE bucket, like the CUDA trainer's flat gradient
world_used = shard.allreduce_bucket(flat)
flat.mul_(1.0 / world_used)
if rank == 0:
    out = shard.unflatten_bucket(flat, [grads[n] for n in names])
    worst = 0.0
    for n, a in zip(names, out):
        ref = torch.from_numpy(t["g:" + n])
        worst = max(worst, (float((a - ref).abs().max()) - 1e-7) / (float(ref.abs().max()) + 1e-12))     # 1e-7: tensors whose gradient is ~0 (key biases)
    print("WORST %.3e" % worst)
dist.destroy_process_group()
'''


def test_data_parallel_bucket_equals_full_batch_gradient(tmp_path):
    """Two gloo ranks each run the step on one of the fixture's two segments; the summed flat bucket / world equals the
    reference's full-batch gradient (every loss term is a mean over positions, shards are equal)."""
    w = tmp_path / "worker.py"
    w.write_text(WORKER)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1", "--master-port", "29533",
           str(w), ROOT]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("WORST")]
    assert line and float(line[0].split()[1]) <= 2e-4, r.stdout[-500:]

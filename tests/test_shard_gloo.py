"""N > 1 host logic on CPU: world_size-2 gloo run of the sharding helpers (the data path itself has no collective)."""
import os
import socket
import subprocess
import sys

from nylon_amt_b200 import shard

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, json
sys.path.insert(0, %r)
import torch, torch.distributed as dist
from nylon_amt_b200 import shard
dist.init_process_group("gloo")
rank, local, ws = shard.world()
assert ws == dist.get_world_size() == 2 and rank == dist.get_rank()
lo, hi = shard.partition(1201, ws, rank)                 # config 3: 1 200 clips (+1 to make it ragged)
mine = torch.zeros(1201, dtype=torch.int32); mine[lo:hi] = 1
dist.all_reduce(mine)
assert bool((mine == 1).all()), "every clip is owned by exactly one rank"
audio, ms = shard.reduce_report(local_audio_s=(hi - lo) * 300.0, local_ms=100.0 + 50.0 * rank)
assert audio == 1201 * 300.0 and ms == 150.0, (audio, ms)
if rank == 0:
    print(json.dumps({"ok": True, "block0": [lo, hi]}))
dist.destroy_process_group()
'''


def test_partition_covers_everything_once():
    for n in (0, 1, 7, 8, 1200, 1758):
        for ws in (1, 2, 3, 4, 8):
            blocks = [shard.partition(n, ws, r) for r in range(ws)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(ws - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_gloo_run(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(script)]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=240, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    assert '"ok": true' in out.stdout

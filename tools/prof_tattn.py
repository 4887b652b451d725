"""GPU box: times the training attention through hft_train_attention on one of the model's attention shapes (default: encoder, 1024
sequences x 2 heads, 256 x 256), tcgen05 kernels against the fp32 CUDA-core kernels.  Also the command that `ncu -k regex:tattn` wraps."""
import argparse
import ctypes
import sys

import torch

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from nylon_amt_b200 import _lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--S", type=int, default=1024)
ap.add_argument("--L", type=int, default=256)
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--tc", type=int, default=1)
ap.add_argument("--drop", type=float, default=0.0)
a = ap.parse_args()
H, heads, dh = 64, 2, 32
S, Lq = a.S, a.L
L = _lib.lib()
qkv = torch.randn(S * Lq, 3 * H, device="cuda")
dctx = torch.randn(S * Lq, H, device="cuda") * 1e-5
ctx = torch.empty(S * Lq, H, device="cuda")
lse = torch.empty(S, heads, Lq, device="cuda")
dqkv = torch.empty(S * Lq, 3 * H, device="cuda")
dbuf = torch.empty(S, heads, Lq, device="cuda")
p = lambda t, off=0: ctypes.c_void_p(t.data_ptr() + 4 * off)
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def run(bwd):
    _lib.check(L.hft_train_attention(a.tc, dh, heads, p(qkv), 3 * H, Lq * 3 * H, p(qkv, H), p(qkv, 2 * H), 3 * H, S, Lq, Lq, a.drop, 7, 1, p(ctx), p(lse),
                                     p(dctx) if bwd else None, p(dqkv), 3 * H, p(dqkv, H), p(dqkv, 2 * H), 3 * H, p(dbuf), st), "hft_train_attention")


for bwd in (False, True):
    for _ in range(3):
        run(bwd)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(a.iters):
        run(bwd)
    e1.record()
    torch.cuda.synchronize()
    print("tc=%d S=%d L=%d drop=%.2f %s: %.1f us per call" % (a.tc, S, Lq, a.drop, "fwd+bwd" if bwd else "fwd", e0.elapsed_time(e1) / a.iters * 1e3))

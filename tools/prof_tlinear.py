"""GPU box: times the Linear kernels of the training step through hft_train_linear / hft_train_linear_wgrad (tcgen05 against fp32 CUDA cores)
on one shape; also the command that `ncu -k regex:tgemm|tdw` wraps.  usage: prof_tlinear.py --form fwd|bwd|wgrad --M .. --N .. --K .. [--tc 0|1]"""
import argparse
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nylon_amt_b200 import _lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--form", default="wgrad")
ap.add_argument("--M", type=int, default=262144)
ap.add_argument("--N", type=int, default=64)
ap.add_argument("--K", type=int, default=64)
ap.add_argument("--tc", type=int, default=1)
ap.add_argument("--iters", type=int, default=10)
a = ap.parse_args()
L = _lib.lib()
M, N, K = a.M, a.N, a.K
p = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
if a.form == "wgrad":
    dy, x, dw, db = torch.randn(M, N, device="cuda") * 1e-6, torch.randn(M, K, device="cuda"), torch.zeros(N, K, device="cuda"), torch.zeros(N, device="cuda")
    run = lambda: _lib.check(L.hft_train_linear_wgrad(a.tc, p(dy), N, p(x), K, p(dw), K, p(db), M, N, K, st), "wgrad")
    nbytes = 4 * M * (N + K)
else:
    x = torch.randn(M, K, device="cuda")
    w = torch.randn(N, K, device="cuda") * 0.1 if a.form == "fwd" else torch.randn(K, N, device="cuda") * 0.1
    b = torch.randn(N, device="cuda") if a.form == "fwd" else None
    c = torch.empty(M, N, device="cuda")
    run = lambda: _lib.check(L.hft_train_linear(a.tc, 0 if a.form == "fwd" else 1, p(x), K, p(w), w.shape[1], p(b), p(c), N, M, N, K, 0, 0, None, 0, 1.0, st), "linear")
    nbytes = 4 * M * (N + K)
for _ in range(3):
    run()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record()
for _ in range(a.iters):
    run()
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) / a.iters * 1e3
print("%s tc=%d M=%d N=%d K=%d: %.1f us per call, %.0f GB/s of algorithmic traffic" % (a.form, a.tc, M, N, K, us, nbytes / us / 1e3))

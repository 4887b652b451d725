"""GPU: the attention of the training step (hft_train_attention: forward with the row log-sum-exp + backward) -- the tcgen05 kernels
(split-fp16 products, tc_train_attn.cuh) and the fp32 CUDA-core kernels -- against torch autograd in fp64 on the same operands
(MultiHeadAttentionLayer.forward, model_spec2midi.py:342-348, with the dropout multipliers the library itself reports through
hft_dropout_mask).  Shapes are the four attention shapes of the model: encoder (256 x 256), decoder cross (88 x 256, with and
without the shared layer-zero queries), decoder self (88 x 88), time (128 x 128)."""
import ctypes

import pytest
import torch

import nylon_amt_b200 as hft
from nylon_amt_b200 import _lib

pytestmark = pytest.mark.gpu

DH, HEADS = 32, 2
H = DH * HEADS


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _run(use_tc, q, ldq, qss, k, v, ldkv, S, Lq, Lk, p_drop, seed, site, d_ctx):
    L = _lib.lib()
    ctx = torch.full((S * Lq, H), float("nan"), device="cuda")
    lse = torch.full((S, HEADS, Lq), float("nan"), device="cuda")
    dq = torch.full((S * Lq, H), float("nan"), device="cuda")
    dk = torch.full((S * Lk, H), float("nan"), device="cuda")
    dv = torch.full((S * Lk, H), float("nan"), device="cuda")
    dbuf = torch.empty((S, HEADS, Lq), device="cuda")
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(L.hft_train_attention(use_tc, DH, HEADS, _ptr(q), ldq, qss, _ptr(k), _ptr(v), ldkv, S, Lq, Lk, p_drop, seed, site, _ptr(ctx), _ptr(lse),
                                     _ptr(d_ctx), _ptr(dq), H, _ptr(dk), _ptr(dv), H, _ptr(dbuf), st), "hft_train_attention")
    torch.cuda.synchronize()
    return ctx, lse, dq, dk, dv


def _reference(qm, km, vm, S, Lq, Lk, mask, d_ctx, shared_q):
    """fp64 autograd: qm [S or 1, Lq, H], km / vm [S, Lk, H], mask [S, HEADS, Lq, Lk] multipliers, d_ctx [S, Lq, H]"""
    q = qm.double().requires_grad_(True)
    k = km.double().requires_grad_(True)
    v = vm.double().requires_grad_(True)
    qh = (q.expand(S, Lq, H) if shared_q else q).reshape(S, Lq, HEADS, DH).permute(0, 2, 1, 3)
    kh = k.reshape(S, Lk, HEADS, DH).permute(0, 2, 1, 3)
    vh = v.reshape(S, Lk, HEADS, DH).permute(0, 2, 1, 3)
    e = qh @ kh.transpose(-1, -2) / (DH ** 0.5)
    lse = torch.logsumexp(e, -1)
    p = torch.softmax(e, -1) * mask.double()
    ctx = (p @ vh).permute(0, 2, 1, 3).reshape(S, Lq, H)
    ctx.backward(d_ctx.double())
    return ctx.detach(), lse.detach(), q.grad, k.grad, v.grad


CASES = [  # (S, Lq, Lk, layout, p_drop)
    (5, 256, 256, "self", 0.0), (3, 256, 256, "self", 0.1), (4, 88, 256, "cross", 0.0), (4, 88, 256, "shared", 0.1), (3, 88, 88, "self", 0.1),
    (3, 128, 128, "self", 0.0), (2, 37, 200, "cross", 0.1), (2, 130, 70, "cross", 0.0),
]


@pytest.mark.parametrize("S,Lq,Lk,layout,p_drop", CASES)
def test_training_attention_matches_autograd(S, Lq, Lk, layout, p_drop):
    g = torch.Generator(device="cuda").manual_seed(1000 + S * 7 + Lq + Lk)
    rnd = lambda *shape: torch.randn(*shape, device="cuda", generator=g)
    if layout == "self":                                  # fused Q | K | V rows, as the encoder / time / decoder-self layers keep them
        qkv = rnd(S * Lq, 3 * H) * 1.5
        q, k, v, ldq, qss, ldkv = qkv, qkv[:, H:], qkv[:, 2 * H:], 3 * H, Lq * 3 * H, 3 * H
        qm, km, vm = qkv[:, :H].reshape(S, Lq, H), qkv[:, H:2 * H].reshape(S, Lk, H), qkv[:, 2 * H:].reshape(S, Lk, H)
    else:
        shared = layout == "shared"
        qb = rnd((1 if shared else S) * Lq, H) * 1.5
        kv = rnd(S * Lk, 2 * H) * 1.5
        q, k, v, ldq, qss, ldkv = qb, kv, kv[:, H:], H, (0 if shared else Lq * H), 2 * H
        qm, km, vm = qb.reshape(-1, Lq, H), kv[:, :H].reshape(S, Lk, H), kv[:, H:].reshape(S, Lk, H)
    d_ctx = rnd(S * Lq, H) * 3e-6                          # gradients of a mean over 90 112 positions are this small
    seed, site = 4321, 5
    n = S * HEADS * Lq * Lk
    mask = torch.ones(n, device="cuda")
    if p_drop > 0:
        _lib.check(_lib.lib().hft_dropout_mask(p_drop, seed, site, n, _ptr(mask), None), "hft_dropout_mask")
    mask = mask.reshape(S, HEADS, Lq, Lk)
    ref = _reference(qm, km, vm, S, Lq, Lk, mask, d_ctx.reshape(S, Lq, H), layout == "shared")
    names = ("ctx", "lse", "dq", "dk", "dv")
    errs = {}
    for use_tc in (1, 0):
        got = _run(use_tc, q, ldq, qss, k, v, ldkv, S, Lq, Lk, p_drop, seed, site, d_ctx)
        for nm, a, b in zip(names, got, ref):
            if nm == "dq" and layout == "shared":         # per-sequence contributions; the reference summed them over the sequences
                a = a.reshape(S, Lq, H).double().sum(0, keepdim=True)
            a = a.double().reshape(b.shape)
            assert torch.isfinite(a).all(), (nm, use_tc)
            errs[(nm, use_tc)] = float((a - b).abs().max() / b.abs().max())
    print("relative-to-max errors (tcgen05 | fp32 CUDA cores):", {nm: (errs[(nm, 1)], errs[(nm, 0)]) for nm in names})
    for nm in names:
        # fp32 class: the split-fp16 products carry 22 mantissa bits; exp2 / the row sums are fp32 in both implementations
        assert errs[(nm, 1)] <= 2e-5, (nm, errs[(nm, 1)], errs[(nm, 0)])
        assert errs[(nm, 0)] <= 2e-5, (nm, errs[(nm, 0)])

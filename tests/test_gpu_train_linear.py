"""GPU: the Linear kernels of the training step (hft_train_linear) -- the tcgen05 kernel (split-fp16 products, tc_train_gemm.cuh) and the
fp32 CUDA-core kernels -- against fp64 matmuls on the same operands: forward y = x W^T + b (+ReLU), input gradient dx = dy W (+ReLU mask,
+accumulate), on the model's shapes (K = 64 / 128, N = 64 / 128 / 144 / 192), ragged row counts and gradient-sized inputs."""
import ctypes

import pytest
import torch

from nylon_amt_b200 import _lib

pytestmark = pytest.mark.gpu


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _linear(use_tc, w_kn, a, lda, w, ldw, bias, c, ldc, m, n, k, relu, accum, mask, ldm, mask_scale):
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(_lib.lib().hft_train_linear(use_tc, w_kn, _ptr(a), lda, _ptr(w), ldw, _ptr(bias), _ptr(c), ldc, m, n, k, relu, accum, _ptr(mask), ldm,
                                           mask_scale, st), "hft_train_linear")
    torch.cuda.synchronize()


CASES = [  # (M, N, K, form, relu_or_mask, accum, input scale)
    (40000, 192, 64, "fwd", 0, 0, 1.0), (12345, 64, 64, "fwd", 0, 1, 1.0), (5000, 128, 64, "fwd", 1, 0, 1.0), (9999, 64, 128, "fwd", 0, 0, 1.0),
    (3000, 144, 64, "fwd", 0, 0, 1.0), (40000, 64, 192 // 3, "bwd", 0, 0, 1e-6), (7777, 64, 128, "bwd", 0, 1, 1e-6), (6000, 128, 64, "bwd", 1, 0, 1e-6),
    (100, 64, 64, "bwd", 0, 0, 1e-6), (129, 16, 64, "fwd", 0, 0, 1.0),
]


@pytest.mark.parametrize("M,N,K,form,rm,accum,scale", CASES)
def test_training_linear_matches_fp64(M, N, K, form, rm, accum, scale):
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    rnd = lambda *shape: torch.randn(*shape, device="cuda", generator=g)
    lda, ldc = K + 64, N + 32                              # rows live inside wider tensors (fused Q | K | V buffers)
    a_full = rnd(M, lda) * scale
    a_full[::7] *= 1e-3                                    # rows of very different magnitude
    a = a_full[:, :K]
    w = rnd(N, K) * 0.1 if form == "fwd" else rnd(K, N) * 0.1
    bias = rnd(N) if form == "fwd" else None
    mask = rnd(M, N) if (form == "bwd" and rm) else None
    c0 = rnd(M, ldc) * scale
    ref = a.double() @ (w.double().t() if form == "fwd" else w.double())
    if form == "fwd":
        ref = ref + bias.double()
        if rm:
            ref = ref.clamp_min(0)
    elif rm:
        ref = torch.where(mask > 0, ref * 1.25, torch.zeros_like(ref))
    if accum:
        ref = ref + c0[:, :N].double()
    errs = {}
    for use_tc in (1, 0):
        c = c0.clone()
        _linear(use_tc, 0 if form == "fwd" else 1, a_full, lda, w, w.shape[1], bias, c, ldc, M, N, K, rm if form == "fwd" else 0, accum, mask, N, 1.25)
        assert torch.equal(c[:, N:], c0[:, N:]), "columns beyond N were touched"
        errs[use_tc] = float((c[:, :N].double() - ref).abs().max() / ref.abs().max())
    print("relative-to-max error: tcgen05 %.3g | fp32 CUDA cores %.3g" % (errs[1], errs[0]))
    assert errs[1] <= 2e-6 and errs[0] <= 2e-6, errs


WG_CASES = [  # (M, N, K)
    (50000, 192, 64), (30001, 64, 64), (20000, 128, 64), (20000, 64, 128), (9000, 144, 64), (300, 64, 64), (70000, 128, 128), (4097, 8, 64),
    # many 32-row stages per CTA: every ring of the kernel wraps several times (a missing proxy fence showed only from 6 stages per CTA on)
    (37888, 192, 64), (56832, 64, 64), (94720, 128, 64), (262144, 64, 64), (262144, 64, 128),
]


@pytest.mark.parametrize("M,N,K", WG_CASES)
def test_training_linear_weight_gradient_matches_fp64(M, N, K):
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    rnd = lambda *shape: torch.randn(*shape, device="cuda", generator=g)
    ldy, ldx = N + 64, K + 32
    dy_full, x_full = rnd(M, ldy) * 2e-6, rnd(M, ldx) * 1.3
    dy_full[::5] *= 1e-2                                   # rows of very different magnitude
    dy_full[: M // 3] *= 30.0
    dw0, db0 = rnd(N, K) * 1e-4, rnd(N) * 1e-4             # the kernels accumulate into what is there
    ref_w = dw0.double() + dy_full[:, :N].double().t() @ x_full[:, :K].double()
    ref_b = db0.double() + dy_full[:, :N].double().sum(0)
    errs = {}
    for use_tc in (1, 0):
        dw, db = dw0.clone(), db0.clone()
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        _lib.check(_lib.lib().hft_train_linear_wgrad(use_tc, _ptr(dy_full), ldy, _ptr(x_full), ldx, _ptr(dw), K, _ptr(db), M, N, K, st), "hft_train_linear_wgrad")
        torch.cuda.synchronize()
        errs[use_tc] = (float((dw.double() - ref_w).abs().max() / ref_w.abs().max()), float((db.double() - ref_b).abs().max() / ref_b.abs().max()))
    print("relative-to-max errors (dW, db): tcgen05 %s | fp32 CUDA cores %s" % (errs[1], errs[0]))
    for use_tc in (1, 0):
        assert errs[use_tc][0] <= 1e-5 and errs[use_tc][1] <= 1e-5, errs   # (the accumulators of a CTA round towards zero: ~4e-6 at 262 144 rows)

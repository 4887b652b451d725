"""GPU parity of the tcgen05 building blocks (through the C ABI component entry points) against fp32 torch math on
the same 16-bit operands, and of the full tensor-core forward against the CPU oracle."""
import ctypes
import math
import os

import numpy as np
import pytest
import torch

import nylon_amt_b200 as hft
from nylon_amt_b200 import _lib
from oracle import hft_oracle as ho

pytestmark = pytest.mark.gpu
DT = {"bf16": torch.bfloat16, "fp16": torch.float16}
EPS = {"bf16": 2.0 ** -8, "fp16": 2.0 ** -11}       # one rounding of the stored 16-bit result


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def tc_linear(kind, epi, a, w, bias, resid=None, gamma=None, beta=None):
    M, K = a.shape
    N = w.shape[0]
    out = torch.empty((M, N), device=a.device, dtype=a.dtype)
    rc = _lib.lib().hft_tc_linear(1 if kind == "bf16" else 0, epi, _ptr(a), _ptr(w), _ptr(bias), M, N, K, _ptr(out), _ptr(resid), _ptr(gamma),
                                  _ptr(beta), None)
    _lib.check(rc, "hft_tc_linear")
    torch.cuda.synchronize()
    return out


@pytest.mark.parametrize("kind", ["bf16", "fp16"])
@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (256, 256, 256), (384, 768, 256), (256, 192, 64), (128, 128, 128), (256, 256, 512), (256, 512, 256)])
def test_linear_store_and_relu(kind, M, N, K):
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = torch.randn((M, K), device="cuda", generator=g).to(DT[kind])
    w = (torch.randn((N, K), device="cuda", generator=g) / math.sqrt(K)).to(DT[kind])
    bias = torch.randn(N, device="cuda", generator=g)
    ref = a.float() @ w.float().t() + bias
    for epi, r in ((0, ref), (1, torch.relu(ref))):
        out = tc_linear(kind, epi, a, w, bias).float()
        err = (out - r).abs().max().item()
        assert err <= EPS[kind] * r.abs().max().item() + 1e-5, (kind, M, N, K, epi, err)


@pytest.mark.parametrize("kind", ["bf16", "fp16"])
@pytest.mark.parametrize("M,N,K", [(256, 256, 256), (128, 64, 64), (256, 256, 512), (128, 64, 128), (128, 128, 128), (384, 128, 256)])
def test_linear_residual_layernorm(kind, M, N, K):
    g = torch.Generator(device="cuda").manual_seed(7 * M + N + K)
    a = torch.randn((M, K), device="cuda", generator=g).to(DT[kind])
    w = (torch.randn((N, K), device="cuda", generator=g) / math.sqrt(K)).to(DT[kind])
    bias = torch.randn(N, device="cuda", generator=g)
    resid = (3 * torch.randn((M, N), device="cuda", generator=g) + 1.5).to(DT[kind])
    gamma = 1 + 0.1 * torch.randn(N, device="cuda", generator=g)
    beta = 0.1 * torch.randn(N, device="cuda", generator=g)
    ref = torch.nn.functional.layer_norm(a.float() @ w.float().t() + bias + resid.float(), (N,), gamma, beta, 1e-5)
    out = tc_linear(kind, 2, a, w, bias, resid, gamma, beta).float()
    err = (out - ref).abs().max().item()
    assert err <= EPS[kind] * ref.abs().max().item() + 1e-4, (kind, M, N, K, err)


@pytest.mark.parametrize("kind", ["bf16", "fp16"])
@pytest.mark.parametrize("M", [256, 512, 256 * 75, 256 * 149])
def test_fused_ffn(kind, M):
    """hft_tc_ffn: LayerNorm(x + fc_2(relu(fc_1(x)))) on CTA pairs with the hidden activation kept in TMEM."""
    H, P = 256, 512
    g = torch.Generator(device="cuda").manual_seed(M)
    x = (torch.randn((M, H), device="cuda", generator=g)).to(DT[kind])
    w1 = (torch.randn((P, H), device="cuda", generator=g) / 16).to(DT[kind])
    w2 = (torch.randn((H, P), device="cuda", generator=g) / 22).to(DT[kind])
    b1 = 0.1 * torch.randn(P, device="cuda", generator=g)
    b2 = 0.1 * torch.randn(H, device="cuda", generator=g)
    gamma = 1 + 0.1 * torch.randn(H, device="cuda", generator=g)
    beta = 0.1 * torch.randn(H, device="cuda", generator=g)
    out = torch.zeros((M, H), device="cuda", dtype=DT[kind])
    rc = _lib.lib().hft_tc_ffn(1 if kind == "bf16" else 0, _ptr(x), _ptr(w1), _ptr(b1), _ptr(w2), _ptr(b2), _ptr(gamma), _ptr(beta), M, _ptr(out), None)
    _lib.check(rc, "hft_tc_ffn")
    torch.cuda.synchronize()
    hid = torch.relu(x.float() @ w1.float().t() + b1).to(DT[kind]).float()          # the kernel rounds the hidden activation to 16 bits
    ref = torch.nn.functional.layer_norm(x.float() + hid @ w2.float().t() + b2, (H,), gamma, beta, 1e-5)
    err = (out.float() - ref).abs().max().item()
    assert err <= 2 * EPS[kind] * ref.abs().max().item() + 1e-3, (kind, M, err)


@pytest.mark.parametrize("kind", ["bf16", "fp16"])
@pytest.mark.parametrize("dh,heads,L,n_seq,probs", [(64, 4, 256, 3, True), (64, 4, 256, 2, False), (64, 4, 128, 5, False), (64, 4, 88, 7, False),
                                                    (64, 4, 256, 80, False), (64, 4, 128, 100, False), (64, 1, 88, 301, False), (64, 2, 256, 1, False),
                                                    (32, 2, 256, 3, True), (32, 2, 128, 3, False), (32, 2, 88, 5, False)])
def test_attention(kind, dh, heads, L, n_seq, probs):
    H = dh * heads
    g = torch.Generator(device="cuda").manual_seed(dh + L + n_seq)
    qkv = (1.5 * torch.randn((n_seq * L, 3 * H), device="cuda", generator=g)).to(DT[kind])
    # rows after the last sequence are read by the 128-row / 96-row boxes when L = 88: keep real memory behind them
    ctx = torch.zeros((n_seq * L, H), device="cuda", dtype=DT[kind])
    pr = torch.zeros((n_seq, heads, L, L), device="cuda") if probs else None
    rc = _lib.lib().hft_tc_attention(1 if kind == "bf16" else 0, dh, heads, _ptr(qkv), n_seq, L, _ptr(ctx), _ptr(pr), None)
    _lib.check(rc, "hft_tc_attention")
    torch.cuda.synchronize()
    x = qkv.float().view(n_seq, L, 3, heads, dh)
    q, k, v = x[:, :, 0].permute(0, 2, 1, 3), x[:, :, 1].permute(0, 2, 1, 3), x[:, :, 2].permute(0, 2, 1, 3)
    att = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(dh), dim=-1)
    ref = (att @ v).permute(0, 2, 1, 3).reshape(n_seq * L, H)
    err = (ctx.float() - ref).abs().max().item()
    # P is rounded to 16 bits before the PV product and the context is stored in 16 bits
    assert err <= 3 * EPS[kind] * max(1.0, v.abs().max().item()), (kind, dh, L, err)
    if probs:
        perr = (pr - att).abs().max().item()
        assert perr <= 2e-5, perr
        assert float((pr.sum(-1) - 1).abs().max()) < 1e-5


@pytest.mark.parametrize("size", ["reduced", "paper"])
def test_forward_split_fp16_meets_fp32_budget(golden_dir, size):
    """HFT_PREC_F16X3: split fp16 operands on tensor cores must meet the north_star fp32 budget (2e-3 abs on every head
    output) against the reference goldens -- the same check the CUDA-core fp32 path passes."""
    from test_gpu_forward import _golden_check, TOL_FP32
    hid, pf, L, h = {"reduced": (64, 128, 2, 2), "paper": (256, 512, 3, 4)}[size]
    g = np.load(os.path.join(golden_dir, "hft_%s.npz" % size))
    model = hft.build_model(hft.default_config(), hid, pf, L, h, seed=1234, device="cuda")
    model.precision = "fp16x3"
    out = model(torch.from_numpy(g["spec"]).cuda())
    worst = _golden_check(out, g, TOL_FP32)
    print("fp16x3", size, worst)


@pytest.mark.parametrize("kind,size", [("fp16", "reduced"), ("bf16", "reduced"), ("fp16", "paper"), ("bf16", "paper")])
def test_forward_single_product_modes_vs_reference_golden(golden_dir, kind, size):
    """The single-product tensor-core modes against the REFERENCE goldens on seeded (freshly initialised) weights.  These weights amplify
    upstream rounding ~20x in the time stack, so only part of the outputs can meet the 16-bit budget (2e-2) here -- the mode
    that meets it everywhere on this fixture is `mixed` (test below), and on trained-like weights fp16 / bf16 meet it on every output
    (tests/test_gpu_trained.py).  Asserted: every output finite; the outputs that do meet 2e-2 on this fixture keep meeting it (A-head
    probabilities, the returned attention, fp16's velocity A and B-head probabilities); the rest is bounded by 1.5x the error the CPU
    emulation of the same rounding points predicts (tools/precision_study.py), as a regression guard, and printed."""
    hid, pf, L, h = {"reduced": (64, 128, 2, 2), "paper": (256, 512, 3, 4)}[size]
    g = np.load(os.path.join(golden_dir, "hft_%s.npz" % size))
    model = hft.build_model(hft.default_config(), hid, pf, L, h, seed=1234, device="cuda")
    model.precision = kind
    out = model(torch.from_numpy(g["spec"]).cuda())
    names = ["onset_A", "offset_A", "mpe_A", "velocity_A", "attention", "onset_B", "offset_B", "mpe_B", "velocity_B"]
    err = {}
    for i, n in enumerate(names):
        o = out[i].cpu()
        assert torch.isfinite(o).all(), n
        if n.startswith("velocity"):
            ref, mine = g[n + "_sub"], o[:, ::8, ::8, :].numpy()
        elif n == "attention":
            ref, mine = g["attention_sub"], o[:, ::16, :, ::11, :].numpy()
        else:
            ref, mine = g[n], o.numpy()
        err[n] = float(np.abs(mine - ref).max())
    print(kind, size, {k: "%.1e" % v for k, v in err.items()})
    assert max(err[n] for n in ("onset_A", "offset_A", "mpe_A", "attention")) <= 2e-2, err
    sig_b = max(err[n] for n in ("onset_B", "offset_B", "mpe_B"))
    if kind == "fp16":
        assert err["velocity_A"] <= 2e-2 and sig_b <= 2e-2, err
        assert err["velocity_B"] <= 0.15, err                  # emulation: 0.10 (paper), measured 0.088
    else:
        assert err["velocity_A"] <= 6e-2 and sig_b <= 6e-2, err      # emulation: 4.0e-2 / 4e-2, measured 3.7e-2 / 3.8e-2
        assert err["velocity_B"] <= 0.45, err                  # emulation: 0.31, measured 0.29


def _forward_in_subprocess(golden_dir, env, out_path, strided):
    """Run the paper-size fp16x3 forward in a fresh process (the library reads its experiment switches once per process)."""
    import subprocess, sys
    code = (
        "import sys, os, numpy as np, torch\n"
        "sys.path.insert(0, %r)\n"
        "import nylon_amt_b200 as hft\n"
        "g = np.load(os.path.join(%r, 'hft_paper.npz'))\n"
        "model = hft.build_model(hft.default_config(), 256, 512, 3, 4, seed=1234, device='cuda')\n"
        "model.precision = 'fp16x3'\n"
        "spec = torch.from_numpy(g['spec']).cuda()\n"
        "if %r:\n"
        "    # the layout AMT.transcript hands over: windows of a [T, 256] feature, strides (128 * 256, 1, 256)\n"
        "    B, NB, W = spec.shape\n"
        "    feat = torch.zeros(128 * (B - 1) + W, NB, device='cuda')\n"
        "    spec0 = spec[0].t().contiguous()\n"
        "    feat[:W] = spec0\n"
        "    spec = feat.as_strided((B, NB, W), (128 * NB, 1, NB))\n"
        "out = model(spec)\n"
        "np.savez(%r, *[t.float().cpu().numpy() for t in out])\n"
    ) % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), golden_dir, strided, out_path)
    e = dict(os.environ)
    e.update(env)
    subprocess.run([sys.executable, "-c", code], check=True, env=e, timeout=300)
    z = np.load(out_path)
    return [z[k] for k in z.files]


@pytest.mark.parametrize("strided", [False, True])
def test_front_tensor_core_kernel_equals_cuda_core_filter(golden_dir, tmp_path, strided):
    """front_tc_kernel (Toeplitz tiles on tcgen05, three split products) against the fp32 CUDA-core filter it replaces
    (HFT_TC_FRONT=0), through the whole paper-size forward: contiguous [B, 256, 192] input and the strided window view of a
    [T, 256] feature that AMT.transcript passes (reference amt.py:88-89)."""
    a = _forward_in_subprocess(golden_dir, {"HFT_TC_FRONT": "1"}, str(tmp_path / "tc.npz"), strided)
    b = _forward_in_subprocess(golden_dir, {"HFT_TC_FRONT": "0"}, str(tmp_path / "cc.npz"), strided)
    worst = max(float(np.abs(x - y).max()) for x, y in zip(a, b))
    print("front tc vs cuda-core, strided=%s: worst abs diff %.3e" % (strided, worst))
    assert all(np.isfinite(x).all() for x in a)
    assert worst <= 1e-3, worst


_SWITCH_BASELINE = {}


@pytest.mark.parametrize("switch", ["HFT_TC_FFN=0", "HFT_TC_PAIR=0", "HFT_TC_ATTN2=0", "HFT_TC_ATTN_PROBS=0", "HFT_TC_WRES=0", "HFT_TC_STAGE_ROWS=16",
                                    "HFT_TC_STAGE_ROWS=8", "HFT_TC_WSTAGES=6"])
def test_experiment_switches_select_equivalent_kernels(golden_dir, tmp_path, switch):
    """Every alternative kernel kept behind an experiment switch ('Experiment switches') computes the same paper-size
    fp16x3 forward as the default path: within 1e-3 abs of it on every output, and inside the 2e-3 budget against the golden."""
    from test_gpu_forward import _golden_check, TOL_FP32
    if "base" not in _SWITCH_BASELINE:
        _SWITCH_BASELINE["base"] = _forward_in_subprocess(golden_dir, {}, str(tmp_path / "base.npz"), False)
    base = _SWITCH_BASELINE["base"]
    k, v = switch.split("=")
    got = _forward_in_subprocess(golden_dir, {k: v}, str(tmp_path / "alt.npz"), False)
    worst = max(float(np.abs(x - y).max()) for x, y in zip(got, base))
    print(switch, "worst abs diff to the default path %.3e" % worst)
    assert worst <= 1e-3, (switch, worst)
    g = np.load(os.path.join(golden_dir, "hft_paper.npz"))
    _golden_check([torch.from_numpy(x) for x in got], g, TOL_FP32)


TOL_16BIT = 2e-2     # north_star: head logits within 2e-2 abs in the 16-bit mode


def test_forward_mixed_meets_16bit_budget_on_every_output(golden_dir):
    """HFT_PREC_MIXED (the 16-bit-class mode: per-GEMM plan of one or three split-fp16 products, include/hft_sm100.h) against the
    REFERENCE goldens (not the repo's own fp32 path): every one of the nine outputs within 2e-2 abs, B heads included."""
    from test_gpu_forward import _golden_check
    g = np.load(os.path.join(golden_dir, "hft_paper.npz"))
    model = hft.build_model(hft.default_config(), 256, 512, 3, 4, seed=1234, device="cuda")
    model.precision = "mixed"
    out = model(torch.from_numpy(g["spec"]).cuda())
    worst = _golden_check(out, g, TOL_16BIT)
    print("mixed paper", worst)
    assert (out[3].argmax(3).cpu().numpy()[:, ::8, ::8] == g["velocity_A_sub"].argmax(3)).mean() > 0.99
    # the mode is real: it differs from fp16x3 and is not worse than the budget anywhere
    model.precision = "fp16x3"
    x3 = model(torch.from_numpy(g["spec"]).cuda())
    assert any(not torch.equal(a, b) for a, b in zip(out, x3))

"""GPU: the CUDA training step (hft_train_forward_backward + hft_adam_step through nylon_amt_b200.training) against the
fixture written by the unmodified reference (tests/golden/train_reduced.npz) and against the CPU oracle."""
import os

import numpy as np
import pytest
import torch

import nylon_amt_b200 as hft
from oracle import train_oracle

pytestmark = pytest.mark.gpu


def _model(golden_dir, dropout=0.0):
    g = np.load(os.path.join(golden_dir, "hft_reduced.npz"))
    model = hft.build_model(hft.default_config(), 64, 128, 2, 2, dropout=dropout, device="cuda")
    model.load_state_dict({k[2:]: torch.from_numpy(g[k]).clone() for k in g.files if k.startswith("w:")})
    return model


def _batch(t):
    return (torch.from_numpy(t["spec"]).cuda(), torch.from_numpy(t["label_onset"]).cuda(), torch.from_numpy(t["label_offset"]).cuda(),
            torch.from_numpy(t["label_mpe"]).cuda(), torch.from_numpy(t["label_velocity"]).cuda())


def test_loss_and_every_gradient_match_reference(golden_dir):
    t = np.load(os.path.join(golden_dir, "train_reduced.npz"))
    model = _model(golden_dir)
    opt = hft.training.Adam(model, lr=float(t["lr"]), batch_size=2)
    loss = opt.forward_backward(*_batch(t))
    torch.cuda.synchronize()
    assert abs(float(loss.item()) - float(t["loss"])) <= 2e-5 * abs(float(t["loss"])), (float(loss.item()), float(t["loss"]))
    worst = {}
    gmax = max(float(np.abs(t[k]).max()) for k in t.files if k.startswith("g:"))
    for name in model._handle().names:
        ref = torch.from_numpy(t["g:" + name])
        got = opt.grad_of(name).cpu()
        err = float((got - ref).abs().max())
        # fp32 summation order over up to 65 536 rows: 2e-4 of the tensor's own scale + 1e-5 of the largest gradient of the model
        # (the floor covers tensors whose true gradient is ~0, e.g. key biases, and small embedding tables)
        tol = 2e-4 * float(ref.abs().max()) + 1e-5 * gmax
        worst[name] = (err, tol)
    bad = {k: v for k, v in worst.items() if v[0] > v[1]}
    assert not bad, bad


def test_adam_step_and_second_iteration_match_reference(golden_dir):
    t = np.load(os.path.join(golden_dir, "train_reduced.npz"))
    model = _model(golden_dir)
    opt = hft.training.Adam(model, lr=float(t["lr"]), batch_size=2)
    batch = _batch(t)
    loss1 = float(hft.training.train_step(model, opt, *batch).item())
    opt.sync_to_module()
    sd = dict(model.named_parameters())
    gmax = max(float(np.abs(t[k]).max()) for k in t.files if k.startswith("g:"))
    for name in model._handle().names:
        ref, g = torch.from_numpy(t["p1:" + name]), torch.from_numpy(t["g:" + name])
        got = sd[name].detach().cpu()
        # Adam's first update is -lr * g / (|g| + eps) = -+lr: exact where |g| stands clear of the summation-order noise
        # (1e-5 * gmax, see the gradient test); where g ~ 0 its sign is noise and the update may differ by up to 2 lr
        firm = g.abs() > 1e-3 * gmax
        assert float((got - ref)[firm].abs().max() if firm.any() else 0.0) <= 2e-7, name
        assert float((got - ref).abs().max()) <= 2.1e-4, name
    loss2 = float(hft.training.train_step(model, opt, *batch).item())
    assert abs(loss1 - float(t["loss"])) <= 2e-5 * abs(float(t["loss"]))
    assert abs(loss2 - float(t["loss2"])) <= 1e-4 * abs(float(t["loss2"])), (loss2, float(t["loss2"]))
    # the inference forward sees the trained weights (hft_model_refresh): outputs moved, and equal the module's own after sync
    model.eval()
    out = model(batch[0])
    assert torch.isfinite(out[5]).all()


def test_gradients_match_cpu_oracle_on_other_data(golden_dir):
    """Different seed / labels / loss weights than the fixture: CUDA step vs the autograd restatement."""
    g = np.load(os.path.join(golden_dir, "hft_reduced.npz"))
    model = _model(golden_dir)
    sd = {k[2:]: torch.from_numpy(g[k]).clone() for k in g.files if k.startswith("w:")}
    spec = torch.from_numpy(g["spec"][:1]).clone()
    spec = spec + 0.05 * torch.randn(spec.shape, generator=torch.Generator().manual_seed(3))
    lab = train_oracle.synthetic_labels(1, seed=11)
    ref_loss, ref_g = train_oracle.loss_and_grads(sd, 2, spec, *lab, weight_A=0.7, weight_B=1.3)
    opt = hft.training.Adam(model, batch_size=1)
    loss = opt.forward_backward(spec.cuda(), *[x.cuda() for x in lab], weight_A=0.7, weight_B=1.3)
    assert abs(float(loss.item()) - ref_loss) <= 2e-5 * abs(ref_loss)
    gmax = max(float(v.abs().max()) for v in ref_g.values())
    for name, ref in ref_g.items():
        err = float((opt.grad_of(name).cpu() - ref).abs().max())
        assert err <= 2e-4 * float(ref.abs().max()) + 1e-5 * gmax, (name, err)


def test_loss_decreases_and_dropout_is_refused(golden_dir):
    t = np.load(os.path.join(golden_dir, "train_reduced.npz"))
    model = _model(golden_dir)
    opt = hft.training.Adam(model, lr=1e-3, batch_size=2)
    batch = _batch(t)
    losses = [float(hft.training.train_step(model, opt, *batch).item()) for _ in range(6)]
    assert losses[-1] < losses[0] - 0.5, losses


def test_dropout_mask_restatement_and_rate():
    """The library's counter-based dropout multiplier equals the numpy restatement bit for bit; keep rate ~ 1 - p."""
    import ctypes
    from nylon_amt_b200 import _lib
    n = 1 << 20
    for p, seed, site in ((0.1, 12345, 0), (0.1, 999, 17), (0.5, 7, 3)):
        out = torch.empty(n, device="cuda")
        _lib.check(_lib.lib().hft_dropout_mask(p, seed, site, n, ctypes.c_void_p(out.data_ptr()), None), "hft_dropout_mask")
        ref = train_oracle.dropout_multiplier(p, seed, site, n)
        assert torch.equal(out.cpu(), ref), (p, seed, site)
        keep = float((ref > 0).float().mean())
        assert abs(keep - (1 - p)) < 3e-3, keep


def test_dropout_training_step_matches_oracle_with_the_same_masks(golden_dir):
    """Train-mode step with p = 0.1 (the reference's setting): loss and every gradient against the autograd restatement that
    applies the same masks at the same sites (embedding, attention probabilities, sub-layer outputs, FFN hidden)."""
    g = np.load(os.path.join(golden_dir, "hft_reduced.npz"))
    model = _model(golden_dir, dropout=0.1)
    sd = {k[2:]: torch.from_numpy(g[k]).clone() for k in g.files if k.startswith("w:")}
    spec = torch.from_numpy(g["spec"][:1]).clone()
    lab = train_oracle.synthetic_labels(1, seed=21)
    opt = hft.training.Adam(model, batch_size=1, seed=5)
    loss = opt.forward_backward(spec.cuda(), *[x.cuda() for x in lab])
    got_loss = float(loss.item())
    # with masks the fp32 evaluation itself sits ~1e-3 of a tensor's scale away from an fp64 evaluation (zeroed + rescaled paths), so the
    # CUDA gradients are compared with the fp64 restatement and must be as close to it as the fp32 restatement is (x3), see the paper-size test
    _, g32 = train_oracle.loss_and_grads_dropout(sd, 2, spec, *lab, p=0.1, seed=opt.last_dropout_seed)
    ref_loss, g64 = train_oracle.loss_and_grads_dropout(sd, 2, spec, *lab, p=0.1, seed=opt.last_dropout_seed, dtype=torch.float64)
    assert abs(got_loss - ref_loss) <= 2e-5 * abs(ref_loss), (got_loss, ref_loss)
    gmax = max(float(v.abs().max()) for v in g64.values())
    bad = {}
    for name, ref in g64.items():
        noise = float((g32[name].double() - ref).abs().max())
        err = float((opt.grad_of(name).cpu().double() - ref).abs().max())
        if err > 3 * noise + 2e-4 * float(ref.abs().max()) + 1e-5 * gmax:
            bad[name] = (err, noise, float(ref.abs().max()))
    assert not bad, bad
    # a second call draws different masks
    loss2 = float(opt.forward_backward(spec.cuda(), *[x.cuda() for x in lab]).item())
    assert loss2 != got_loss


def test_paper_size_gradients_match_cpu_oracle(golden_dir):
    """Paper-size model (hid 256, ff 512, 3+3 layers, 4 heads: head_dim 64 kernels), one segment.  With seeded xavier weights this
    model amplifies rounding ~10x per stack: the reference's own fp32 arithmetic deviates from an fp64 evaluation by
    up to 1.2e-3 of a tensor's scale.  So the CUDA gradients are compared with the fp64 restatement, and must be as close to it as
    the fp32 restatement is (x3), on top of the reduced-model rule."""
    model = hft.build_model(hft.default_config(), 256, 512, 3, 4, dropout=0.0, seed=1234, device="cuda")
    sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    g = np.load(os.path.join(golden_dir, "hft_reduced.npz"))
    spec = torch.from_numpy(g["spec"][:1]).clone()
    lab = train_oracle.synthetic_labels(1, seed=5)
    _, g32 = train_oracle.loss_and_grads(sd, 4, spec, *lab)
    ref_loss, g64 = train_oracle.loss_and_grads(sd, 4, spec, *lab, dtype=torch.float64)
    opt = hft.training.Adam(model, batch_size=1)
    loss = opt.forward_backward(spec.cuda(), *[x.cuda() for x in lab])
    assert abs(float(loss.item()) - ref_loss) <= 2e-5 * abs(ref_loss), (float(loss.item()), ref_loss)
    gmax = max(float(v.abs().max()) for v in g64.values())
    bad = {}
    for name, ref in g64.items():
        noise = float((g32[name].double() - ref).abs().max())
        err = float((opt.grad_of(name).cpu().double() - ref).abs().max())
        if err > 3 * noise + 2e-4 * float(ref.abs().max()) + 1e-5 * gmax:
            bad[name] = (err, noise, float(ref.abs().max()))
    assert not bad, bad


# ---- the reference's own loop against the mirror module: model.train(); model(x); criteria; loss.backward(); stock optimiser ----------------

def _reference_loop_loss(model, spec, on, off, mpe, vel, wA=1.0, wB=1.0):
    """train.py:89-151 restated: forward in train mode, the 8 criteria the reference builds (m_training.py:150-160), weighted sum."""
    bce, ce = torch.nn.BCELoss(), torch.nn.CrossEntropyLoss()
    oA, fA, mA, vA, attention, oB, fB, mB, vB = model(spec)
    assert attention is None
    lA = bce(oA.contiguous().view(-1), on.view(-1)) + bce(fA.contiguous().view(-1), off.view(-1)) + bce(mA.contiguous().view(-1), mpe.view(-1)) + \
        ce(vA.contiguous().view(-1, vA.shape[-1]), vel.view(-1))
    lB = bce(oB.contiguous().view(-1), on.view(-1)) + bce(fB.contiguous().view(-1), off.view(-1)) + bce(mB.contiguous().view(-1), mpe.view(-1)) + \
        ce(vB.contiguous().view(-1, vB.shape[-1]), vel.view(-1))
    return wA * lA + wB * lB


def test_train_mode_forward_autograd_and_stock_adam_match_reference(golden_dir):
    """The UNMODIFIED reference loop (train.py:89-158) against nylon_amt_b200.Model_SPEC2MIDI: train-mode forward (hft_train_forward) ->
    torch criteria -> loss.backward() (hft_train_backward behind one autograd node) -> torch.optim.Adam.step(); loss, all 115 gradients,
    the parameters after the step and the loss of the second iteration against the fixture the reference wrote."""
    t = np.load(os.path.join(golden_dir, "train_reduced.npz"))
    model = _model(golden_dir).train()
    optimizer = torch.optim.Adam(model.parameters(), lr=float(t["lr"]))
    batch = _batch(t)
    optimizer.zero_grad()
    loss = _reference_loop_loss(model, *batch)
    assert abs(float(loss.item()) - float(t["loss"])) <= 2e-5 * abs(float(t["loss"])), (float(loss.item()), float(t["loss"]))
    loss.backward()
    gmax = max(float(np.abs(t[k]).max()) for k in t.files if k.startswith("g:"))
    bad = {}
    for name, p in model.named_parameters():
        ref = torch.from_numpy(t["g:" + name])
        assert p.grad is not None, name
        err = float((p.grad.cpu() - ref).abs().max())
        if err > 2e-4 * float(ref.abs().max()) + 1e-5 * gmax:
            bad[name] = err
    assert not bad, bad
    optimizer.step()
    for name, p in model.named_parameters():
        ref, g = torch.from_numpy(t["p1:" + name]), torch.from_numpy(t["g:" + name])
        firm = g.abs() > 1e-3 * gmax
        assert float((p.detach().cpu() - ref)[firm].abs().max() if firm.any() else 0.0) <= 2e-7, name
    optimizer.zero_grad()
    loss2 = _reference_loop_loss(model, *batch)          # the forward re-uploads the parameters the stock optimiser changed in place
    assert abs(float(loss2.item()) - float(t["loss2"])) <= 1e-4 * abs(float(t["loss2"])), (float(loss2.item()), float(t["loss2"]))


def test_train_mode_forward_outputs_equal_eval_forward_without_dropout(golden_dir):
    g = np.load(os.path.join(golden_dir, "hft_reduced.npz"))
    model = _model(golden_dir)
    spec = torch.from_numpy(g["spec"]).cuda()
    model.eval()
    model.precision = "fp32"
    ev = [x.clone() for x in model(spec)]
    model.train()
    with torch.no_grad():
        tr = model(spec)
    for i in (0, 1, 2, 3, 5, 6, 7, 8):
        assert float((tr[i] - ev[i]).abs().max()) <= 2e-4, i
    # with p = 0.1 the outputs move, and differ from call to call
    md = _model(golden_dir, dropout=0.1).train()
    a = md(spec)[5].detach().clone()
    b = md(spec)[5].detach()
    assert float((a - ev[5]).abs().max()) > 1e-4 and not torch.equal(a, b)


def test_partial_last_batch(golden_dir):
    """DataLoader(drop_last=False): the last batch of an epoch is smaller than the trainer's capacity; result = a trainer of that size."""
    t = np.load(os.path.join(golden_dir, "train_reduced.npz"))
    batch = [x[:1] for x in _batch(t)]
    big = hft.training.Adam(_model(golden_dir), batch_size=3)
    one = hft.training.Adam(_model(golden_dir), batch_size=1)
    lb = float(big.forward_backward(*batch).item())
    lo = float(one.forward_backward(*batch).item())
    assert abs(lb - lo) <= 1e-6 * abs(lo), (lb, lo)                            # fp32 atomics: the summation order differs from run to run
    gmax = float(one.grads.abs().max())
    assert float((big.grads - one.grads).abs().max()) <= 1e-5 * gmax
    with pytest.raises(RuntimeError):
        big.forward_backward(*[torch.cat([x, x, x, x]) for x in batch])


def test_adam_is_a_torch_optimizer_with_checkpoint_round_trip(golden_dir, tmp_path):
    """m_training.py:147 ReduceLROnPlateau(optimizer); :373-384 pickle.dump(model) / state_dict / optimizer_dict; :277-278 resume."""
    import pickle
    t = np.load(os.path.join(golden_dir, "train_reduced.npz"))
    batch = _batch(t)
    model = _model(golden_dir)
    opt = hft.training.Adam(model, lr=1e-3, batch_size=2)
    assert isinstance(opt, torch.optim.Optimizer)
    sched = torch.optim.lr_scheduler.ReduceLROnPlateau(opt, mode="min", factor=0.5, patience=0)
    for _ in range(2):
        hft.training.train_step(model, opt, *batch)
    sched.step(1.0); sched.step(2.0)                       # a plateau: lr halves, and the next step uses it
    assert abs(opt.lr - 5e-4) < 1e-12
    # checkpoint: the module's state_dict sees the trained weights without an explicit sync
    sd_model = {k: v.detach().clone() for k, v in model.state_dict().items()}
    flat = opt._flat_params()
    for name, (off, numel) in opt._slot.items():
        assert torch.equal(sd_model[name].reshape(-1), flat[off:off + numel]), name
    blob = pickle.dumps(model)
    assert torch.equal(pickle.loads(blob).state_dict()["decoder_spec2midi.fc_onset_time.bias"].cpu(), sd_model["decoder_spec2midi.fc_onset_time.bias"].cpu())
    sd_opt = opt.state_dict()
    torch.save({"model_dict": sd_model, "optimizer_dict": sd_opt, "scheduler_dict": sched.state_dict()}, str(tmp_path / "ck.pt"))
    l3 = float(hft.training.train_step(model, opt, *batch).item())
    # resume in a fresh module + optimiser: step 3 reproduces
    ck = torch.load(str(tmp_path / "ck.pt"), weights_only=False)
    model2 = _model(golden_dir)
    model2.load_state_dict(ck["model_dict"])
    opt2 = hft.training.Adam(model2, lr=1e-3, batch_size=2)
    opt2.load_state_dict(ck["optimizer_dict"])
    assert abs(opt2.lr - 5e-4) < 1e-12 and opt2.step_count == 2
    l3b = float(hft.training.train_step(model2, opt2, *batch).item())
    assert abs(l3b - l3) <= 1e-6 * abs(l3), (l3, l3b)                      # fp32 atomics in the loss / dW sums: run-to-run order noise
    opt.sync_to_module(); opt2.sync_to_module()
    for (n, a), (_, b) in zip(model.named_parameters(), model2.named_parameters()):
        # the step is lr * m / (sqrt(v) + eps) with lr = 5e-4: resumed and uninterrupted runs may differ by summation-order noise only
        assert float((a.detach() - b.detach()).abs().max()) <= 2e-5, n
    # the same optimizer_dict loads into a stock torch.optim.Adam over the same parameters
    stock = torch.optim.Adam(model2.parameters(), lr=1e-3)
    stock.load_state_dict(ck["optimizer_dict"])
    assert len(stock.state) == len(list(model2.parameters()))


def test_train_returns_with_module_synced(golden_dir):
    t = np.load(os.path.join(golden_dir, "train_reduced.npz"))
    model = _model(golden_dir)
    before = model.decoder_spec2midi.fc_onset_time.bias.detach().clone()
    opt = hft.training.Adam(model, lr=1e-3, batch_size=2)
    batch = _batch(t)
    it = [tuple(x[:2] for x in batch), tuple(x[:1] for x in batch)]          # second batch is partial
    hft.training.train(model, it, opt)
    after = model.decoder_spec2midi.fc_onset_time.bias.detach()
    assert not torch.equal(before, after) and not opt._stale


def test_unmodified_reference_train_and_valid_functions_drive_the_mirror(golden_dir):
    """The reference's OWN train() and valid() (hftt_code/training/train.py:63-160, :168-262, staged unmodified in oracle/_ref by
    __graft_entry__.build()) called with the mirror module, stock criteria and a stock torch.optim.Adam -- nothing of this repo's
    training.py on the path.  The fixture batch twice = train.py's epoch over a two-item iterator: epoch loss = (loss + loss2) / 2 of
    tests/golden/train_reduced.npz."""
    import importlib
    ref_copy = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref")
    if not os.path.isfile(os.path.join(ref_copy, "hftt_code", "training", "train.py")):
        pytest.skip("oracle/_ref is not staged (run __graft_entry__.build() where /root/reference exists)")
    old = os.environ.get("NYLON_REF_ROOT")
    os.environ["NYLON_REF_ROOT"] = ref_copy
    try:
        from oracle import _refload
        importlib.reload(_refload)
        ref_train = _refload.load_train()
    finally:
        if old is None:
            os.environ.pop("NYLON_REF_ROOT", None)
        else:
            os.environ["NYLON_REF_ROOT"] = old
    t = np.load(os.path.join(golden_dir, "train_reduced.npz"))
    model = _model(golden_dir)
    batch = tuple(x.cpu() for x in _batch(t))                     # train() moves every item to `device` itself (train.py:73-77)
    iterator = [batch, batch]
    optimizer = torch.optim.Adam(model.parameters(), lr=float(t["lr"]))
    crit = [torch.nn.BCELoss(), torch.nn.BCELoss(), torch.nn.BCELoss(), torch.nn.CrossEntropyLoss()] * 2
    epoch_loss = ref_train.train(model, iterator, optimizer, *crit, 1.0, 1.0, "cuda", False)
    want = 0.5 * (float(t["loss"]) + float(t["loss2"]))
    assert abs(epoch_loss - want) <= 1e-4 * abs(want), (epoch_loss, want)
    assert model.training
    total, n = ref_train.valid(model, iterator, *crit, 1.0, 1.0, "cuda", False)
    assert n == 2 and not model.training and np.isfinite(total) and total / n < float(t["loss"])      # two Adam steps later the loss is lower


def test_backward_through_an_overwritten_tape_fails_loudly(golden_dir):
    """One activation tape per module: a second train-mode forward overwrites it, so the first graph's backward must raise, not
    return gradients of the wrong forward."""
    t = np.load(os.path.join(golden_dir, "train_reduced.npz"))
    model = _model(golden_dir).train()
    batch = _batch(t)
    loss1 = _reference_loop_loss(model, *batch)
    loss2 = _reference_loop_loss(model, *batch)
    with pytest.raises(RuntimeError, match="overwritten"):
        loss1.backward()
    loss2.backward()                                   # the latest forward still owns the tape
    assert all(p.grad is not None for p in model.parameters())
    with torch.no_grad():                              # no graph requested: forward only
        out = model(batch[0])
    assert out[4] is None and not out[0].requires_grad

"""Training step on the B200 path -- mirror of hftt_code/training/train.py:63-160 (`train`) and of the optimiser the
reference builds in hftt_code/training/m_training.py:146 (`optim.Adam(model.parameters(), lr=...)`).

Two ways to run the step, both on libhft_sm100.so (include/hft_sm100.h):

  * unmodified reference loop: `model.train(); out = model(input_spec); loss = criteria(out...); loss.backward(); optimizer.step()` with a
    stock `torch.optim.Adam` -- Model_SPEC2MIDI.forward in train mode runs hft_train_forward and hands autograd a node whose backward is
    hft_train_backward (nylon_amt_b200/model_spec2midi.py).
  * fused step: `opt = hft.training.Adam(model, lr=1e-4, batch_size=8)`; `hft.training.train_step(model, opt, spec, onset, offset, mpe, velocity)`:
    forward, the 8-term loss, backward (hft_train_forward_backward) and Adam on the flat parameter vector (hft_adam_step) without
    materialising the outputs; in the data-parallel configuration the ONE flat gradient bucket is all-reduced over NCCL between backward and
    the Adam step.

`Adam` is a torch.optim.Optimizer: `param_groups[0]['lr']` is read every step (so `ReduceLROnPlateau(optimizer)`, m_training.py:147, works)
and `state_dict()` / `load_state_dict()` use torch.optim.Adam's own layout (`optimizer_dict` of m_training.py:382 round-trips, also into a
stock torch.optim.Adam).  Dropout: the p of the model's nn.Dropout modules (the reference builds them with 0.1) is applied with the library's
counter-based masks (hft_trainer_set_dropout, fresh seed every step); p = 0 gives the deterministic parity configuration.
"""
import ctypes
import weakref

import torch
import torch.nn as nn

from . import _lib


class Adam(torch.optim.Optimizer):
    """torch.optim.Adam(params, lr, betas, eps) semantics (no weight decay / amsgrad) on the library's flat parameter vector."""

    def __init__(self, model, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, batch_size=8, process_group=None, seed=0):
        ps = sorted({float(m.p) for m in model.modules() if isinstance(m, nn.Dropout)})
        if len(ps) > 1:
            raise NotImplementedError("the B200 training step applies ONE dropout probability to every site (the reference does too); got %r" % (ps,))
        named = list(model.named_parameters())
        super().__init__([p for _, p in named], dict(lr=float(lr), betas=(float(betas[0]), float(betas[1])), eps=float(eps), weight_decay=0,
                                                    amsgrad=False, maximize=False, foreach=None, capturable=False, differentiable=False, fused=None))
        self.p_drop = ps[0] if ps else 0.0
        self.model, self.batch_size, self.group, self.step_count = model, int(batch_size), process_group, 0
        h = model.sync_weights()
        L = _lib.lib()
        self.n = int(L.hft_model_param_floats(h.ptr))
        self.offsets = [int(L.hft_model_param_offset(h.ptr, i)) for i in range(len(h.names))]
        self._slot = {name: (self.offsets[i], h.numel[i]) for i, name in enumerate(h.names)}
        self._names = [n for n, _ in named]                  # param_groups[0]['params'] order = model.parameters() order
        dev = next(model.parameters()).device
        self.device = dev
        self.grads = torch.zeros(self.n, device=dev)
        self.exp_avg = torch.zeros(self.n, device=dev)
        self.exp_avg_sq = torch.zeros(self.n, device=dev)
        self.loss = torch.zeros(1, device=dev)
        self.trainer = ctypes.c_void_p()
        self._stale = False
        with torch.cuda.device(dev):
            _lib.check(L.hft_trainer_create(ctypes.byref(self.trainer), h.ptr, self.batch_size), "hft_trainer_create")
        import torch.distributed as dist
        self._rank = 0
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(process_group) > 1:
            # data parallel: every replica starts from rank 0's parameters and draws its own dropout masks
            self._rank = dist.get_rank(process_group)
            flat = self._flat_params()
            dist.broadcast(flat, src=dist.get_global_rank(process_group, 0) if process_group is not None else 0, group=process_group)
            self._write_flat_params(flat)
        self.seed = int(seed) * 1000003 + self._rank
        # state_dict() / pickling of the module must see the trained weights, which live in the library's arena between steps:
        # Model_SPEC2MIDI.state_dict / __getstate__ look this reference up and call sync_if_stale()
        model.__dict__["_hft_trained_by"] = weakref.ref(self)

    def __del__(self):
        try:
            if getattr(self, "trainer", None) is not None and self.trainer.value:
                _lib.lib().hft_trainer_destroy(self.trainer)
        except Exception:
            pass

    # ---- torch.optim.Adam-compatible hyper-parameters and checkpoint format ------------------------------------------------------------
    @property
    def lr(self):
        return float(self.param_groups[0]["lr"])

    @lr.setter
    def lr(self, v):
        self.param_groups[0]["lr"] = float(v)

    @property
    def betas(self):
        return tuple(float(b) for b in self.param_groups[0]["betas"])

    @property
    def eps(self):
        return float(self.param_groups[0]["eps"])

    def state_dict(self):
        """torch.optim.Adam's layout: state[i] = {step, exp_avg, exp_avg_sq} for parameter i of model.parameters() (m_training.py:382)."""
        state = {}
        if self.step_count > 0:
            shapes = dict(self.model.named_parameters())
            for i, name in enumerate(self._names):
                off, numel = self._slot[name]
                state[i] = {"step": torch.tensor(float(self.step_count)),
                            "exp_avg": self.exp_avg[off:off + numel].view(shapes[name].shape).clone(),
                            "exp_avg_sq": self.exp_avg_sq[off:off + numel].view(shapes[name].shape).clone()}
        groups = [{k: v for k, v in g.items() if k != "params"} for g in self.param_groups]
        groups[0]["params"] = list(range(len(self._names)))
        return {"state": state, "param_groups": groups}

    def load_state_dict(self, sd):
        """Accepts a state_dict of this class or of a stock torch.optim.Adam over the same model.parameters() (m_training.py:277)."""
        g = sd["param_groups"][0]
        if len(g["params"]) != len(self._names):
            raise ValueError("optimizer state_dict has %d parameters, the model has %d" % (len(g["params"]), len(self._names)))
        for k in ("lr", "betas", "eps"):
            if k in g:
                self.param_groups[0][k] = g[k]
        self.exp_avg.zero_()
        self.exp_avg_sq.zero_()
        steps = set()
        for i, st in sd["state"].items():
            off, numel = self._slot[self._names[int(i)]]
            self.exp_avg[off:off + numel].copy_(st["exp_avg"].reshape(-1))
            self.exp_avg_sq[off:off + numel].copy_(st["exp_avg_sq"].reshape(-1))
            steps.add(int(float(st["step"])))
        if len(steps) > 1:
            raise ValueError("per-parameter step counts differ (%r): the flat Adam keeps one" % (sorted(steps),))
        self.step_count = steps.pop() if steps else 0

    # ---- views ---------------------------------------------------------------------------------------------------
    def grad_of(self, name):
        """Gradient of one state_dict entry as a view into the flat bucket."""
        off, numel = self._slot[name]
        p = dict(self.model.named_parameters())[name]
        return self.grads[off:off + numel].view(p.shape)

    def zero_grad(self, set_to_none=True):          # train.py:104 -- hft_train_forward_backward overwrites the bucket, nothing to do
        pass

    def _flat_params(self):
        h = self.model._handle()
        flat = torch.empty(self.n, device=self.device)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().hft_model_get_params(h.ptr, ctypes.c_void_p(flat.data_ptr()), ctypes.c_void_p(stream)), "hft_model_get_params")
        return flat

    def _write_flat_params(self, flat):
        """flat vector -> the module's nn.Parameters -> re-registered with the library."""
        sd = dict(self.model.named_parameters())
        with torch.no_grad():
            for name, (off, numel) in self._slot.items():
                sd[name].copy_(flat[off:off + numel].view(sd[name].shape))
        self.model.sync_weights(force=True)

    # ---- the three phases of train.py:105-158 ----------------------------------------------------------------------
    def forward_backward(self, input_spec, label_onset, label_offset, label_mpe, label_velocity, weight_A=1.0, weight_B=1.0):
        m = self.model
        m.sync_weights()                   # no-op between steps: the module's tensors are untouched, the library's arena holds the current parameters
        e = m.encoder_spec2midi
        x = input_spec if input_spec.dtype == torch.float32 else input_spec.float()
        if not x.is_cuda:
            raise RuntimeError("input_spec is on %s: the B200 path has no CPU fallback" % x.device)
        B = x.shape[0] if x.dim() == 3 else -1
        if x.dim() != 3 or not (1 <= B <= self.batch_size) or x.shape[1] != e.n_bin or x.shape[2] != e.n_frame + e.n_proc - 1:
            raise RuntimeError("input_spec must be [B <= %d, %d, %d], got %s" % (self.batch_size, e.n_bin, e.n_frame + e.n_proc - 1, tuple(x.shape)))
        lab = [t.to(self.device, torch.float32).contiguous() for t in (label_onset, label_offset, label_mpe)]
        vel = label_velocity.to(self.device, torch.int64).contiguous()
        for t in lab + [vel]:
            if t.numel() != B * e.n_frame * m.decoder_spec2midi.n_note:
                raise RuntimeError("label tensors must be [B, n_frame, n_note] with the batch size of input_spec")
        stream = torch.cuda.current_stream(self.device).cuda_stream
        with torch.cuda.device(self.device):
            if self.p_drop > 0.0:              # fresh masks every iteration: (seed, iteration) -> 32-bit seed of the counter-based generator
                self._fb_calls = getattr(self, "_fb_calls", 0) + 1
                step_seed = (self.seed * 2654435761 + self._fb_calls * 40503) & 0xFFFFFFFF
                _lib.check(_lib.lib().hft_trainer_set_dropout(self.trainer, self.p_drop, step_seed), "hft_trainer_set_dropout")
                self.last_dropout_seed = step_seed
            _lib.check(_lib.lib().hft_train_forward_backward_n(
                self.trainer, B, ctypes.c_void_p(x.data_ptr()), x.stride(0), x.stride(1), x.stride(2), ctypes.c_void_p(lab[0].data_ptr()),
                ctypes.c_void_p(lab[1].data_ptr()), ctypes.c_void_p(lab[2].data_ptr()), ctypes.c_void_p(vel.data_ptr()), float(weight_A), float(weight_B),
                ctypes.c_void_p(self.loss.data_ptr()), ctypes.c_void_p(self.grads.data_ptr()), ctypes.c_void_p(stream)), "hft_train_forward_backward")
        return self.loss

    def all_reduce(self):
        """Data parallel: one flat-bucket sum over the ranks (NCCL over NVLink on the GPU box, gloo in the CPU tests)."""
        from . import shard
        return shard.allreduce_bucket(self.grads, self.group)

    def step(self, world=1):
        if callable(world):                     # Optimizer.step(closure) convention
            world()
            world = 1
        self.step_count += 1
        h = self.model._handle()
        L = _lib.lib()
        stream = torch.cuda.current_stream(self.device).cuda_stream
        with torch.cuda.device(self.device):
            _lib.check(L.hft_adam_step(L.hft_model_params(h.ptr), ctypes.c_void_p(self.grads.data_ptr()), ctypes.c_void_p(self.exp_avg.data_ptr()),
                                       ctypes.c_void_p(self.exp_avg_sq.data_ptr()), self.n, self.lr, self.betas[0], self.betas[1], self.eps,
                                       self.step_count, 1.0 / world, ctypes.c_void_p(stream)), "hft_adam_step")
            _lib.check(L.hft_model_refresh(h.ptr, ctypes.c_void_p(stream)), "hft_model_refresh")
        self._stale = True                      # the module's nn.Parameters lag behind the library's arena until sync_to_module()

    def sync_if_stale(self):
        if self._stale:
            self.sync_to_module()

    def sync_to_module(self):
        """Copy the library's (trained) flat parameters back into the module's nn.Parameters (for state_dict / checkpoints, m_training.py:275,373).
        Called automatically at the end of train(), by model.state_dict() and when the module is pickled."""
        h = self.model._handle()
        flat = self._flat_params()
        sd = dict(self.model.named_parameters())
        with torch.no_grad():
            for name, (off, numel) in self._slot.items():
                sd[name].copy_(flat[off:off + numel].view(sd[name].shape))
        # the module's tensors now equal the library's: refresh the change stamp so the next forward does not re-upload
        h.stamp = tuple((p.detach().data_ptr(), p.detach()._version) for p in (sd[n] for n in h.names))
        self._stale = False
        return self.model


def train_step(model, optimizer, input_spec, label_onset, label_offset, label_mpe, label_velocity, weight_A=1.0, weight_B=1.0):
    """One iteration of the loop body of train.py:72-158: zero_grad, forward, loss, backward, (all-reduce,) optimizer.step().
    Returns the loss as a device scalar tensor (train.py:159 calls .item() on it)."""
    optimizer.zero_grad()
    loss = optimizer.forward_backward(input_spec, label_onset, label_offset, label_mpe, label_velocity, weight_A, weight_B)
    world = optimizer.all_reduce()
    optimizer.step(world)
    return loss


def train(model, iterator, optimizer, criterion_onset_A=None, criterion_offset_A=None, criterion_mpe_A=None, criterion_velocity_A=None,
          criterion_onset_B=None, criterion_offset_B=None, criterion_mpe_B=None, criterion_velocity_B=None, weight_A=1.0, weight_B=1.0,
          device=None, verbose_flag=False):
    """Same argument list as the reference's train() (train.py:63-68).  The criteria are fixed by the library (BCELoss x6,
    CrossEntropyLoss x2, mean reduction) and only checked here.  Batches may be smaller than the optimizer's batch_size (the last batch of a
    DataLoader with drop_last=False).  On return the module's parameters hold the trained weights (state_dict / pickle.dump are safe)."""
    for c in (criterion_onset_A, criterion_offset_A, criterion_mpe_A, criterion_onset_B, criterion_offset_B, criterion_mpe_B):
        if c is not None and not isinstance(c, nn.BCELoss):
            raise RuntimeError("libhft_sm100 implements BCELoss for onset / offset / mpe")
    for c in (criterion_velocity_A, criterion_velocity_B):
        if c is not None and not isinstance(c, nn.CrossEntropyLoss):
            raise RuntimeError("libhft_sm100 implements CrossEntropyLoss for velocity")
    epoch_loss, n = 0.0, 0
    for input_spec, label_onset, label_offset, label_mpe, label_velocity in iterator:
        loss = train_step(model, optimizer, input_spec.to(optimizer.device, non_blocking=True), label_onset, label_offset, label_mpe, label_velocity,
                          weight_A, weight_B)
        epoch_loss += float(loss.item())
        n += 1
    optimizer.sync_to_module()
    return epoch_loss / max(n, 1)

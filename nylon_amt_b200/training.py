"""Training step on the B200 path -- mirror of hftt_code/training/train.py:63-160 (`train`) and of the optimiser the
reference builds in hftt_code/training/m_training.py:146 (`optim.Adam(model.parameters(), lr=...)`).

    opt = hft.training.Adam(model, lr=1e-4, batch_size=8)            # flat fp32 parameters / gradients / moments on the device
    loss = hft.training.train_step(model, opt, spec, onset, offset, mpe, velocity, weight_A=1.0, weight_B=1.0)

Forward, the 8-term loss, backward and Adam run in libhft_sm100.so (hft_train_forward_backward / hft_adam_step,
include/hft_sm100.h); PyTorch owns the flat gradient / moment tensors, and in the data-parallel configuration it
all-reduces the ONE flat gradient bucket over NCCL between backward and the Adam step.  Dropout: the p of the
model's nn.Dropout modules (the reference builds them with 0.1) is applied with the library's counter-based masks
(hft_trainer_set_dropout, fresh seed every step); p = 0 gives the deterministic parity configuration.
"""
import ctypes

import torch
import torch.nn as nn

from . import _lib


class Adam:
    """torch.optim.Adam(params, lr, betas, eps) semantics (no weight decay / amsgrad) on the library's flat parameter vector."""

    def __init__(self, model, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, batch_size=8, process_group=None, seed=0):
        ps = sorted({float(m.p) for m in model.modules() if isinstance(m, nn.Dropout)})
        if len(ps) > 1:
            raise NotImplementedError("the B200 training step applies ONE dropout probability to every site (the reference does too); got %r" % (ps,))
        self.p_drop = ps[0] if ps else 0.0
        self.seed = int(seed)
        self.model, self.lr, self.betas, self.eps = model, float(lr), (float(betas[0]), float(betas[1])), float(eps)
        self.batch_size, self.group, self.step_count = int(batch_size), process_group, 0
        h = model.sync_weights()
        L = _lib.lib()
        self.n = int(L.hft_model_param_floats(h.ptr))
        self.offsets = [int(L.hft_model_param_offset(h.ptr, i)) for i in range(len(h.names))]
        dev = next(model.parameters()).device
        self.device = dev
        self.grads = torch.zeros(self.n, device=dev)
        self.exp_avg = torch.zeros(self.n, device=dev)
        self.exp_avg_sq = torch.zeros(self.n, device=dev)
        self.loss = torch.zeros(1, device=dev)
        self.trainer = ctypes.c_void_p()
        with torch.cuda.device(dev):
            _lib.check(L.hft_trainer_create(ctypes.byref(self.trainer), h.ptr, self.batch_size), "hft_trainer_create")

    def __del__(self):
        try:
            if getattr(self, "trainer", None) is not None and self.trainer.value:
                _lib.lib().hft_trainer_destroy(self.trainer)
        except Exception:
            pass

    # ---- views ---------------------------------------------------------------------------------------------------
    def grad_of(self, name):
        """Gradient of one state_dict entry as a view into the flat bucket."""
        h = self.model._handle()
        i = h.names.index(name)
        p = dict(self.model.named_parameters())[name]
        return self.grads[self.offsets[i]:self.offsets[i] + h.numel[i]].view(p.shape)

    def zero_grad(self):          # train.py:104 -- hft_train_forward_backward overwrites the bucket, nothing to do
        pass

    # ---- the three phases of train.py:105-158 ----------------------------------------------------------------------
    def forward_backward(self, input_spec, label_onset, label_offset, label_mpe, label_velocity, weight_A=1.0, weight_B=1.0):
        m = self.model
        h = m.sync_weights()
        e = m.encoder_spec2midi
        x = input_spec if input_spec.dtype == torch.float32 else input_spec.float()
        if not x.is_cuda:
            raise RuntimeError("input_spec is on %s: the B200 path has no CPU fallback" % x.device)
        if x.dim() != 3 or x.shape[0] != self.batch_size or x.shape[1] != e.n_bin or x.shape[2] != e.n_frame + e.n_proc - 1:
            raise RuntimeError("input_spec must be [%d, %d, %d], got %s" % (self.batch_size, e.n_bin, e.n_frame + e.n_proc - 1, tuple(x.shape)))
        lab = [t.to(self.device, torch.float32).contiguous() for t in (label_onset, label_offset, label_mpe)]
        vel = label_velocity.to(self.device, torch.int64).contiguous()
        for t in lab + [vel]:
            if t.numel() != self.batch_size * e.n_frame * m.decoder_spec2midi.n_note:
                raise RuntimeError("label tensors must be [B, n_frame, n_note]")
        stream = torch.cuda.current_stream(self.device).cuda_stream
        with torch.cuda.device(self.device):
            if self.p_drop > 0.0:              # fresh masks every iteration: (seed, iteration) -> 32-bit seed of the counter-based generator
                self._fb_calls = getattr(self, "_fb_calls", 0) + 1
                step_seed = (self.seed * 2654435761 + self._fb_calls * 40503) & 0xFFFFFFFF
                _lib.check(_lib.lib().hft_trainer_set_dropout(self.trainer, self.p_drop, step_seed), "hft_trainer_set_dropout")
                self.last_dropout_seed = step_seed
            _lib.check(_lib.lib().hft_train_forward_backward(
                self.trainer, ctypes.c_void_p(x.data_ptr()), x.stride(0), x.stride(1), x.stride(2), ctypes.c_void_p(lab[0].data_ptr()),
                ctypes.c_void_p(lab[1].data_ptr()), ctypes.c_void_p(lab[2].data_ptr()), ctypes.c_void_p(vel.data_ptr()), float(weight_A), float(weight_B),
                ctypes.c_void_p(self.loss.data_ptr()), ctypes.c_void_p(self.grads.data_ptr()), ctypes.c_void_p(stream)), "hft_train_forward_backward")
        return self.loss

    def all_reduce(self):
        """Data parallel: one flat-bucket sum over the ranks (NCCL over NVLink on the GPU box, gloo in the CPU tests)."""
        from . import shard
        return shard.allreduce_bucket(self.grads, self.group)

    def step(self, world=1):
        self.step_count += 1
        h = self.model._handle()
        L = _lib.lib()
        stream = torch.cuda.current_stream(self.device).cuda_stream
        with torch.cuda.device(self.device):
            _lib.check(L.hft_adam_step(L.hft_model_params(h.ptr), ctypes.c_void_p(self.grads.data_ptr()), ctypes.c_void_p(self.exp_avg.data_ptr()),
                                       ctypes.c_void_p(self.exp_avg_sq.data_ptr()), self.n, self.lr, self.betas[0], self.betas[1], self.eps,
                                       self.step_count, 1.0 / world, ctypes.c_void_p(stream)), "hft_adam_step")
            _lib.check(L.hft_model_refresh(h.ptr, ctypes.c_void_p(stream)), "hft_model_refresh")
        self._stale = True

    def sync_to_module(self):
        """Copy the library's (trained) flat parameters back into the module's nn.Parameters (for state_dict / checkpoints, m_training.py:275)."""
        h = self.model._handle()
        L = _lib.lib()
        flat = torch.empty(self.n, device=self.device)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        with torch.cuda.device(self.device):
            _lib.check(L.hft_model_get_params(h.ptr, ctypes.c_void_p(flat.data_ptr()), ctypes.c_void_p(stream)), "hft_model_get_params")
        sd = dict(self.model.named_parameters())
        with torch.no_grad():
            for name, off, numel in zip(h.names, self.offsets, h.numel):
                sd[name].copy_(flat[off:off + numel].view(sd[name].shape))
        # the module's tensors now equal the library's: refresh the change stamp so the next forward does not re-upload
        h.stamp = tuple((p.detach().data_ptr(), p.detach()._version) for p in (sd[n] for n in h.names))
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
    CrossEntropyLoss x2, mean reduction) and only checked here."""
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
    return epoch_loss / max(n, 1)

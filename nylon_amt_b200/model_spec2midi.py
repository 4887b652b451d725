"""Host-side mirror of the reference's hftt_code/model/model_spec2midi.py for the forward hot path.

Same class names, constructor arguments, attribute names and state_dict schema as the reference
(Model_SPEC2MIDI :9, Encoder_SPEC2MIDI :41, Decoder_SPEC2MIDI :112, EncoderLayer :222, DecoderLayer_Zero :247,
DecoderLayer :274, MultiHeadAttentionLayer :308, PositionwiseFeedforwardLayer :362), so that
`load_state_dict(checkpoint['model_dict'])` (hftt_code/training/m_training.py:275) and modules pickled by the
reference (hftt_code/model/amt.py:24-25, after nylon_amt_b200.install_reference_aliases()) keep working.

The sub-modules only hold parameters.  `Model_SPEC2MIDI.forward` does not run PyTorch ops: it hands the
parameters and the input to libhft_sm100.so (hand-written sm_100a CUDA behind the C ABI in include/hft_sm100.h).
There is no CPU path and no eager fallback: parameters or inputs that are not on a CUDA device raise.
"""
import ctypes

import torch
import torch.nn as nn

from . import _lib


class MultiHeadAttentionLayer(nn.Module):
    """Parameter container of model_spec2midi.py:308-320 (fc_q / fc_k / fc_v / fc_o)."""

    def __init__(self, hid_dim, n_heads, dropout, device):
        super().__init__()
        assert hid_dim % n_heads == 0
        self.hid_dim = hid_dim
        self.n_heads = n_heads
        self.head_dim = hid_dim // n_heads
        self.fc_q = nn.Linear(hid_dim, hid_dim)
        self.fc_k = nn.Linear(hid_dim, hid_dim)
        self.fc_v = nn.Linear(hid_dim, hid_dim)
        self.fc_o = nn.Linear(hid_dim, hid_dim)
        self.dropout = nn.Dropout(dropout)


class PositionwiseFeedforwardLayer(nn.Module):
    """Parameter container of model_spec2midi.py:362-367 (fc_1 / fc_2)."""

    def __init__(self, hid_dim, pf_dim, dropout):
        super().__init__()
        self.fc_1 = nn.Linear(hid_dim, pf_dim)
        self.fc_2 = nn.Linear(pf_dim, hid_dim)
        self.dropout = nn.Dropout(dropout)


class EncoderLayer(nn.Module):
    """model_spec2midi.py:222-228: one LayerNorm shared by both residual sites."""

    def __init__(self, hid_dim, n_heads, pf_dim, dropout, device):
        super().__init__()
        self.layer_norm = nn.LayerNorm(hid_dim)
        self.self_attention = MultiHeadAttentionLayer(hid_dim, n_heads, dropout, device)
        self.positionwise_feedforward = PositionwiseFeedforwardLayer(hid_dim, pf_dim, dropout)
        self.dropout = nn.Dropout(dropout)


class DecoderLayer_Zero(nn.Module):
    """model_spec2midi.py:247-253: cross-attention + FFN, no self-attention."""

    def __init__(self, hid_dim, n_heads, pf_dim, dropout, device):
        super().__init__()
        self.layer_norm = nn.LayerNorm(hid_dim)
        self.encoder_attention = MultiHeadAttentionLayer(hid_dim, n_heads, dropout, device)
        self.positionwise_feedforward = PositionwiseFeedforwardLayer(hid_dim, pf_dim, dropout)
        self.dropout = nn.Dropout(dropout)


class DecoderLayer(nn.Module):
    """model_spec2midi.py:274-281: self-attention, cross-attention, FFN; one LayerNorm for the three sites."""

    def __init__(self, hid_dim, n_heads, pf_dim, dropout, device):
        super().__init__()
        self.layer_norm = nn.LayerNorm(hid_dim)
        self.self_attention = MultiHeadAttentionLayer(hid_dim, n_heads, dropout, device)
        self.encoder_attention = MultiHeadAttentionLayer(hid_dim, n_heads, dropout, device)
        self.positionwise_feedforward = PositionwiseFeedforwardLayer(hid_dim, pf_dim, dropout)
        self.dropout = nn.Dropout(dropout)


class Encoder_SPEC2MIDI(nn.Module):
    """Same constructor as model_spec2midi.py:42; parameters conv / tok_embedding_freq / pos_embedding_freq / layers_freq."""

    def __init__(self, n_margin, n_frame, n_bin, cnn_channel, cnn_kernel, hid_dim, n_layers, n_heads, pf_dim, dropout, device):
        super().__init__()
        self.device = device
        self.n_frame = n_frame
        self.n_bin = n_bin
        self.cnn_channel = cnn_channel
        self.cnn_kernel = cnn_kernel
        self.hid_dim = hid_dim
        self.conv = nn.Conv2d(1, self.cnn_channel, kernel_size=(1, self.cnn_kernel))
        self.n_proc = n_margin * 2 + 1
        self.cnn_dim = self.cnn_channel * (self.n_proc - (self.cnn_kernel - 1))
        self.tok_embedding_freq = nn.Linear(self.cnn_dim, hid_dim)
        self.pos_embedding_freq = nn.Embedding(n_bin, hid_dim)
        self.layers_freq = nn.ModuleList([EncoderLayer(hid_dim, n_heads, pf_dim, dropout, device) for _ in range(n_layers)])
        self.dropout = nn.Dropout(dropout)

    def forward(self, spec_in):
        """model_spec2midi.py:60-106: spec_in [B, n_bin, n_frame + 2 margin] -> [B, n_frame, n_bin, hid_dim] (eval semantics, fp32
        CUDA-core kernels through hft_forward_encoder).  The library model holds both halves, so the module must belong to a
        Model_SPEC2MIDI; the fused Model_SPEC2MIDI.forward is the hot path."""
        return _parent_of(self)._forward_encoder(spec_in)


class Decoder_SPEC2MIDI(nn.Module):
    """Same constructor as model_spec2midi.py:113."""

    def __init__(self, n_frame, n_bin, n_note, n_velocity, hid_dim, n_layers, n_heads, pf_dim, dropout, device):
        super().__init__()
        self.device = device
        self.n_note = n_note
        self.n_frame = n_frame
        self.n_velocity = n_velocity
        self.n_bin = n_bin
        self.hid_dim = hid_dim
        self.sigmoid = nn.Sigmoid()
        self.dropout = nn.Dropout(dropout)
        self.pos_embedding_freq = nn.Embedding(n_note, hid_dim)
        self.layer_zero_freq = DecoderLayer_Zero(hid_dim, n_heads, pf_dim, dropout, device)
        self.layers_freq = nn.ModuleList([DecoderLayer(hid_dim, n_heads, pf_dim, dropout, device) for _ in range(n_layers - 1)])
        self.fc_onset_freq = nn.Linear(hid_dim, 1)
        self.fc_offset_freq = nn.Linear(hid_dim, 1)
        self.fc_mpe_freq = nn.Linear(hid_dim, 1)
        self.fc_velocity_freq = nn.Linear(hid_dim, self.n_velocity)
        self.pos_embedding_time = nn.Embedding(n_frame, hid_dim)
        self.layers_time = nn.ModuleList([EncoderLayer(hid_dim, n_heads, pf_dim, dropout, device) for _ in range(n_layers)])
        self.fc_onset_time = nn.Linear(hid_dim, 1)
        self.fc_offset_time = nn.Linear(hid_dim, 1)
        self.fc_mpe_time = nn.Linear(hid_dim, 1)
        self.fc_velocity_time = nn.Linear(hid_dim, self.n_velocity)

    def forward(self, enc_spec):
        """model_spec2midi.py:145-216: enc_spec [B, n_frame, n_bin, hid_dim] -> (onset_A, offset_A, mpe_A, velocity_A, attention,
        onset_B, offset_B, mpe_B, velocity_B) (eval semantics, fp32 CUDA-core kernels through hft_forward_decoder)."""
        return _parent_of(self)._forward_decoder(enc_spec)


def _child_getstate(self):
    st = self.__dict__.copy()
    st.pop("_hft_parent", None)         # a weak reference does not pickle; Model_SPEC2MIDI.__setstate__ restores the link
    return st


Encoder_SPEC2MIDI.__getstate__ = _child_getstate
Decoder_SPEC2MIDI.__getstate__ = _child_getstate


def _parent_of(module):
    ref = module.__dict__.get("_hft_parent")
    parent = ref() if ref is not None else None
    if parent is None:
        raise RuntimeError("%s.forward needs the Model_SPEC2MIDI it belongs to (libhft_sm100 keeps encoder and decoder weights in one "
                           "model handle); construct Model_SPEC2MIDI(encoder, decoder) first" % type(module).__name__)
    return parent


class _Handle:
    """Owns the hft_model handle of one module instance (freed with the module)."""

    def __init__(self, dims):
        self.ptr = ctypes.c_void_p()
        _lib.check(_lib.lib().hft_model_create(ctypes.byref(self.ptr), ctypes.byref(dims)), "hft_model_create")
        n = _lib.lib().hft_model_num_weights(self.ptr)
        self.names = [_lib.lib().hft_model_weight_name(self.ptr, i).decode() for i in range(n)]
        self.numel = [_lib.lib().hft_model_weight_numel(self.ptr, i) for i in range(n)]
        self.stamp = None
        self.max_batch = None

    def __del__(self):
        try:
            if self.ptr:
                _lib.lib().hft_model_destroy(self.ptr)
                self.ptr = None
        except Exception:
            pass


class Model_SPEC2MIDI(nn.Module):
    """Model_SPEC2MIDI(encoder, decoder).forward(input_spec[B,256,192]) -> the 9-tuple of model_spec2midi.py:35."""

    # 'fp16x3': tcgen05 with split fp16 operands (fp32-class: meets the 2e-3 parity budget) -- the default;
    # 'fp32': CUDA-core fp32; 'fp16' / 'bf16': single-product tensor-core modes (faster, lower precision)
    precision = "fp16x3"
    max_batch = 8           # segments processed per internal pass (bounds the workspace)

    def __init__(self, encoder, decoder):
        super().__init__()
        self.encoder_spec2midi = encoder
        self.decoder_spec2midi = decoder
        import weakref
        encoder.__dict__["_hft_parent"] = decoder.__dict__["_hft_parent"] = weakref.ref(self)

    # ---- handle / weights ---------------------------------------------------------------------------------
    def _dims(self):
        e, d = self.encoder_spec2midi, self.decoder_spec2midi
        sa = e.layers_freq[0].self_attention
        dims = _lib.hft_dims()
        dims.n_margin = (e.n_proc - 1) // 2
        dims.n_frame, dims.n_bin = e.n_frame, e.n_bin
        dims.cnn_channel, dims.cnn_kernel = e.cnn_channel, e.cnn_kernel
        dims.hid_dim = e.hid_dim
        dims.pf_dim = e.layers_freq[0].positionwise_feedforward.fc_1.out_features
        dims.n_enc_layers = len(e.layers_freq)
        dims.n_dec_layers = len(d.layers_time)
        dims.n_heads = sa.n_heads
        dims.n_note, dims.n_velocity = d.n_note, d.n_velocity
        if d.layer_zero_freq.encoder_attention.n_heads != sa.n_heads:
            raise RuntimeError("encoder and decoder head counts differ; unsupported by libhft_sm100")
        return dims

    def _handle(self):
        h = self.__dict__.get("_hft")
        if h is None:
            h = _Handle(self._dims())
            self.__dict__["_hft"] = h
        return h

    def sync_weights(self, force=False):
        """(Re-)register the parameters with the library when they changed (load_state_dict, optimizer step, .to())."""
        h = self._handle()
        sd = dict(self.named_parameters())
        tensors = []
        for name, numel in zip(h.names, h.numel):
            if name not in sd:
                raise RuntimeError("parameter %s missing from the module (schema mismatch)" % name)
            p = sd[name]
            if p.numel() != numel:
                raise RuntimeError("parameter %s has %d elements, expected %d" % (name, p.numel(), numel))
            if not p.is_cuda:
                raise RuntimeError("parameter %s is on %s: the B200 path has no CPU fallback, move the model to cuda" % (name, p.device))
            tensors.append(p.detach())
        stamp = tuple((t.data_ptr(), t._version) for t in tensors)
        if force or stamp != h.stamp:
            keep = [t if (t.dtype == torch.float32 and t.is_contiguous()) else t.float().contiguous() for t in tensors]
            arr = (ctypes.c_void_p * len(keep))(*[t.data_ptr() for t in keep])
            stream = torch.cuda.current_stream(keep[0].device).cuda_stream
            with torch.cuda.device(keep[0].device):
                _lib.check(_lib.lib().hft_model_set_weights(h.ptr, arr, len(keep), ctypes.c_void_p(stream)), "hft_model_set_weights")
            h.stamp = stamp
        return h

    def release_workspace(self):
        """Give the library's activation work spaces back (AMT.transcript sizes them for 48 segments per call: 14 GB in fp16x3)."""
        h = self.__dict__.get("_hft")
        if h is not None:
            _lib.check(_lib.lib().hft_model_release_workspace(h.ptr), "hft_model_release_workspace")

    # ---- forward ------------------------------------------------------------------------------------------
    def forward(self, input_spec):
        if self.training:
            return self._forward_train(input_spec)
        if not input_spec.is_cuda:
            raise RuntimeError("input_spec is on %s: the B200 path has no CPU fallback" % input_spec.device)
        e, d = self.encoder_spec2midi, self.decoder_spec2midi
        x = input_spec if input_spec.dtype == torch.float32 else input_spec.float()
        if x.dim() != 3 or x.shape[1] != e.n_bin or x.shape[2] != e.n_frame + e.n_proc - 1:
            raise RuntimeError("input_spec must be [B, %d, %d], got %s" % (e.n_bin, e.n_frame + e.n_proc - 1, tuple(x.shape)))
        h = self.sync_weights()
        if h.max_batch != self.max_batch:
            _lib.check(_lib.lib().hft_model_set_max_batch(h.ptr, int(self.max_batch)), "hft_model_set_max_batch")
            h.max_batch = self.max_batch
        B, F, N, V = x.shape[0], e.n_frame, d.n_note, d.n_velocity
        heads = e.layers_freq[0].self_attention.n_heads
        dev = x.device
        opt = dict(device=dev, dtype=torch.float32)
        outs = [torch.empty((B, F, N), **opt), torch.empty((B, F, N), **opt), torch.empty((B, F, N), **opt),
                torch.empty((B, F, N, V), **opt), torch.empty((B, F, heads, N, e.n_bin), **opt),
                torch.empty((B, F, N), **opt), torch.empty((B, F, N), **opt), torch.empty((B, F, N), **opt),
                torch.empty((B, F, N, V), **opt)]
        o = _lib.hft_outputs(*[ctypes.c_void_p(t.data_ptr()) for t in outs])
        stream = torch.cuda.current_stream(dev).cuda_stream
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().hft_forward(h.ptr, _lib.PREC[self.precision], ctypes.c_void_p(x.data_ptr()), x.stride(0), x.stride(1),
                                              x.stride(2), B, ctypes.byref(o), ctypes.c_void_p(stream)), "hft_forward")
        return tuple(outs)

    # ---- the two halves as separate calls (Encoder_SPEC2MIDI.forward / Decoder_SPEC2MIDI.forward of the reference) -----------------
    def _prepare(self):
        if self.training and any(isinstance(m, nn.Dropout) and m.p > 0 for m in self.modules()):
            raise NotImplementedError("the separate encoder / decoder calls have eval semantics; train through Model_SPEC2MIDI.forward")
        h = self.sync_weights()
        if h.max_batch != self.max_batch:
            _lib.check(_lib.lib().hft_model_set_max_batch(h.ptr, int(self.max_batch)), "hft_model_set_max_batch")
            h.max_batch = self.max_batch
        return h

    def _forward_encoder(self, spec_in):
        x = self._check_input(spec_in)
        h = self._prepare()
        e = self.encoder_spec2midi
        out = torch.empty((x.shape[0], e.n_frame, e.n_bin, e.hid_dim), device=x.device, dtype=torch.float32)
        stream = torch.cuda.current_stream(x.device).cuda_stream
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().hft_forward_encoder(h.ptr, ctypes.c_void_p(x.data_ptr()), x.stride(0), x.stride(1), x.stride(2), x.shape[0],
                                                      ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(stream)), "hft_forward_encoder")
        return out

    def _forward_decoder(self, enc_spec):
        e, d = self.encoder_spec2midi, self.decoder_spec2midi
        if not enc_spec.is_cuda:
            raise RuntimeError("enc_spec is on %s: the B200 path has no CPU fallback" % enc_spec.device)
        if enc_spec.dim() != 4 or tuple(enc_spec.shape[1:]) != (e.n_frame, e.n_bin, e.hid_dim):
            raise RuntimeError("enc_spec must be [B, %d, %d, %d], got %s" % (e.n_frame, e.n_bin, e.hid_dim, tuple(enc_spec.shape)))
        x = enc_spec.float().contiguous()
        h = self._prepare()
        B, F, N, V = x.shape[0], e.n_frame, d.n_note, d.n_velocity
        heads = e.layers_freq[0].self_attention.n_heads
        opt = dict(device=x.device, dtype=torch.float32)
        outs = [torch.empty((B, F, N), **opt) for _ in range(3)] + [torch.empty((B, F, N, V), **opt), torch.empty((B, F, heads, N, e.n_bin), **opt)] + \
               [torch.empty((B, F, N), **opt) for _ in range(3)] + [torch.empty((B, F, N, V), **opt)]
        ptrs = [ctypes.c_void_p(t.data_ptr()) for t in outs]
        o = _lib.hft_outputs(*ptrs, None, None)
        stream = torch.cuda.current_stream(x.device).cuda_stream
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().hft_forward_decoder(h.ptr, ctypes.c_void_p(x.data_ptr()), B, ctypes.byref(o), ctypes.c_void_p(stream)), "hft_forward_decoder")
        return tuple(outs)

    def forward_into(self, input_spec, outs, want_attention=True, velocity_argmax=None):
        """Same computation writing into caller-owned output tensors (no allocation on the hot path).  Entries of `outs`
        may be None (that output is not written).  velocity_argmax: optional pair of int8 tensors [B, n_frame, n_note] that
        receive argmax(velocity logits) of the A / B heads straight from the heads GEMM's epilogue (what AMT.transcript keeps
        of the logits, reference amt.py:107,113)."""
        h = self.sync_weights()
        if h.max_batch != self.max_batch:
            _lib.check(_lib.lib().hft_model_set_max_batch(h.ptr, int(self.max_batch)), "hft_model_set_max_batch")
            h.max_batch = self.max_batch
        x = input_spec
        ptrs = [ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(None) for t in outs]
        if not want_attention:
            ptrs[4] = ctypes.c_void_p(None)
        if velocity_argmax is not None:
            for t in velocity_argmax:
                if t.dtype != torch.int8 or not t.is_cuda or not t.is_contiguous():
                    raise RuntimeError("velocity_argmax tensors must be contiguous CUDA int8")
            ptrs += [ctypes.c_void_p(t.data_ptr()) for t in velocity_argmax]
        o = _lib.hft_outputs(*ptrs)
        stream = torch.cuda.current_stream(x.device).cuda_stream
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().hft_forward(h.ptr, _lib.PREC[self.precision], ctypes.c_void_p(x.data_ptr()), x.stride(0), x.stride(1),
                                              x.stride(2), x.shape[0], ctypes.byref(o), ctypes.c_void_p(stream)), "hft_forward")
        return outs

    # ---- train mode (train.py:89-90: model.train(); outputs = model(input_spec)) ------------------------------------------
    def _check_input(self, input_spec):
        if not input_spec.is_cuda:
            raise RuntimeError("input_spec is on %s: the B200 path has no CPU fallback" % input_spec.device)
        e = self.encoder_spec2midi
        x = input_spec if input_spec.dtype == torch.float32 else input_spec.float()
        if x.dim() != 3 or x.shape[1] != e.n_bin or x.shape[2] != e.n_frame + e.n_proc - 1:
            raise RuntimeError("input_spec must be [B, %d, %d], got %s" % (e.n_bin, e.n_frame + e.n_proc - 1, tuple(x.shape)))
        return x

    def _trainer(self, batch):
        """hft_trainer (activation tape) of this module, grown to `batch` segments on demand."""
        t = self.__dict__.get("_hft_trainer")
        if t is None or t.capacity < batch:
            t = _Trainer(self._handle(), batch, next(self.parameters()).device)
            self.__dict__["_hft_trainer"] = t
        return t

    def _forward_train(self, input_spec):
        """Train-mode forward: dropout active (the p of the module's nn.Dropout layers), fp32 kernels with an activation tape, and -- when
        autograd is recording -- a graph node whose backward is hft_train_backward, so the reference's loop (criteria on the outputs,
        loss.backward(), a stock torch optimiser; train.py:139-158) runs unchanged.  The attention member of the 9-tuple is None in train
        mode (train.py never reads it; the tape keeps row log-sum-exps, not the probabilities)."""
        x = self._check_input(input_spec)
        ps = sorted({float(m.p) for m in self.modules() if isinstance(m, nn.Dropout)})
        if len(ps) > 1:
            raise NotImplementedError("the B200 training step applies ONE dropout probability to every site (the reference does too); got %r" % (ps,))
        p_drop = ps[0] if ps else 0.0
        names = self._handle().names
        sd = dict(self.named_parameters())
        outs = _TrainForward.apply(self, x, p_drop, *[sd[n] for n in names])
        return tuple(outs[:4]) + (None,) + tuple(outs[4:])

    def _sync_trained(self):
        """The fused training step (nylon_amt_b200.training.Adam) keeps the current parameters in the library's arena between steps;
        anything that reads the module's tensors from outside (state_dict, pickle) first copies them back."""
        opt = self.__dict__.get("_hft_trained_by")
        opt = opt() if opt is not None else None
        if opt is not None:
            opt.sync_if_stale()

    def state_dict(self, *args, **kwargs):                # m_training.py:384 torch.save({'model_dict': model.state_dict(), ...})
        self._sync_trained()
        return super().state_dict(*args, **kwargs)

    def __setstate__(self, state):
        super().__setstate__(state)
        import weakref
        for child in (self.encoder_spec2midi, self.decoder_spec2midi):
            child.__dict__["_hft_parent"] = weakref.ref(self)

    def __getstate__(self):
        self._sync_trained()          # pickle.dump(model) after training (m_training.py:373) must see the trained weights
        st = self.__dict__.copy()
        for k in ("_hft", "_hft_trainer", "_hft_trained_by"):   # device handles are rebuilt lazily after unpickling
            st.pop(k, None)
        return st


class _Trainer:
    """Owns one hft_trainer handle (activation tape + gradient work space) and the flat gradient bucket."""

    def __init__(self, handle, capacity, device):
        self.capacity, self.device, self.handle = int(capacity), device, handle
        self.ptr = ctypes.c_void_p()
        L = _lib.lib()
        with torch.cuda.device(device):
            _lib.check(L.hft_trainer_create(ctypes.byref(self.ptr), handle.ptr, self.capacity), "hft_trainer_create")
        self.n = int(L.hft_model_param_floats(handle.ptr))
        self.offsets = [int(L.hft_model_param_offset(handle.ptr, i)) for i in range(len(handle.names))]
        self.grads = torch.zeros(self.n, device=device)
        self.tape_id = 0               # bumped by every train-mode forward: a backward must meet the tape its own forward wrote

    def __del__(self):
        try:
            if self.ptr:
                _lib.lib().hft_trainer_destroy(self.ptr)
                self.ptr = None
        except Exception:
            pass


class _TrainForward(torch.autograd.Function):
    """model(input_spec) in train mode as ONE autograd node: forward = hft_train_forward, backward = hft_train_backward."""

    @staticmethod
    def forward(ctx, model, x, p_drop, *params):
        h = model.sync_weights()
        e, d = model.encoder_spec2midi, model.decoder_spec2midi
        B, F, N, V = x.shape[0], e.n_frame, d.n_note, d.n_velocity
        t = model._trainer(B)
        dev = x.device
        opt = dict(device=dev, dtype=torch.float32)
        outs = [torch.empty((B, F, N), **opt) for _ in range(3)] + [torch.empty((B, F, N, V), **opt)] + \
               [torch.empty((B, F, N), **opt) for _ in range(3)] + [torch.empty((B, F, N, V), **opt)]
        ptrs = [ctypes.c_void_p(o.data_ptr()) for o in outs]
        o = _lib.hft_outputs(ptrs[0], ptrs[1], ptrs[2], ptrs[3], None, ptrs[4], ptrs[5], ptrs[6], ptrs[7], None, None)
        # dropout masks: counter-based, seeded from torch's CPU generator so torch.manual_seed() governs the run
        seed = int(torch.randint(0, 2 ** 31 - 1, (1,)).item()) if p_drop > 0.0 else 0
        stream = torch.cuda.current_stream(dev).cuda_stream
        L = _lib.lib()
        with torch.cuda.device(dev):
            _lib.check(L.hft_trainer_set_dropout(t.ptr, float(p_drop), seed), "hft_trainer_set_dropout")
            _lib.check(L.hft_train_forward(t.ptr, B, ctypes.c_void_p(x.data_ptr()), x.stride(0), x.stride(1), x.stride(2), ctypes.byref(o),
                                           ctypes.c_void_p(stream)), "hft_train_forward")
        ctx.model, ctx.trainer, ctx.x, ctx.handle = model, t, x, h
        t.tape_id += 1
        ctx.tape_id = t.tape_id
        ctx.shapes = [p.shape for p in params]
        return tuple(outs)

    @staticmethod
    def backward(ctx, *gouts):
        t, x = ctx.trainer, ctx.x
        if t.tape_id != ctx.tape_id:
            raise RuntimeError("backward through a train-mode forward whose activation tape was overwritten by a later forward of the same module")
        keep = [None if g is None else (g if (g.dtype == torch.float32 and g.is_contiguous()) else g.float().contiguous()) for g in gouts]
        ptrs = [ctypes.c_void_p(g.data_ptr()) if g is not None else None for g in keep]
        o = _lib.hft_outputs(ptrs[0], ptrs[1], ptrs[2], ptrs[3], None, ptrs[4], ptrs[5], ptrs[6], ptrs[7], None, None)
        stream = torch.cuda.current_stream(x.device).cuda_stream
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().hft_train_backward(t.ptr, ctypes.c_void_p(x.data_ptr()), x.stride(0), x.stride(1), x.stride(2), ctypes.byref(o),
                                                     ctypes.c_void_p(t.grads.data_ptr()), ctypes.c_void_p(stream)), "hft_train_backward")
        grads = [t.grads[off:off + shp.numel()].view(shp).clone() for off, shp in zip(t.offsets, ctx.shapes)]
        return (None, None, None) + tuple(grads)

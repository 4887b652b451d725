"""CPU oracle for the hFT-Transformer forward  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product path (nylon_amt_b200) never does.

A functional fp32 restatement (plain tensor ops on a state_dict, no nn.Module) of
`Model_SPEC2MIDI.forward` (reference hftt_code/model/model_spec2midi.py:15-35):

  encoder front   model_spec2midi.py:63-95    unfold(65) -> Conv2d(1,4,(1,5)) -> Linear(244->H) -> *sqrt(H)+pos
  EncoderLayer    model_spec2midi.py:230-245  x=LN(x+MHA(x)); x=LN(x+FFN(x))     (one shared LayerNorm)
  MHA             model_spec2midi.py:322-360  softmax(QK^T/sqrt(d)) V -> fc_o    (returns probs too)
  FFN             model_spec2midi.py:369-378  fc_2(relu(fc_1(x)))
  DecoderLayer_Zero :255-272, DecoderLayer :283-306 (one shared LayerNorm for 3 sites)
  decoder         model_spec2midi.py:145-216  88 pitch queries, heads A, time re-layout, SAtime, heads B

It is a floating-point path, so the restatement is in torch fp32 on CPU (the "torch fp32 reference"
the task allows for floating-point kernels).  Eval mode: every dropout is the identity.

Parity pin: tests/golden/hft_*.npz hold outputs of the *reference modules themselves*
(/root/reference imported by oracle/make_golden.py in the build container) on seeded weights and
inputs; tests/test_oracle_hft.py checks this restatement against them.

`gemm_in` / `store` let tests/tools emulate reduced-precision GEMM operands and a reduced-precision
activation stream (e.g. bf16 rounding) to study the error budget; the default is exact fp32.
"""
import math

import torch
import torch.nn.functional as F


def dims_from_state_dict(sd):
    """Recover (hid, pf, n_enc_layers, n_dec_layers, cnn_channel, cnn_kernel) from tensor shapes."""
    hid = sd["encoder_spec2midi.tok_embedding_freq.weight"].shape[0]
    pf = sd["encoder_spec2midi.layers_freq.0.positionwise_feedforward.fc_1.weight"].shape[0]
    n_enc = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("encoder_spec2midi.layers_freq."))
    n_dec = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("decoder_spec2midi.layers_time."))
    ch, _, _, kw = sd["encoder_spec2midi.conv.weight"].shape
    return hid, pf, n_enc, n_dec, ch, kw


class Oracle:
    def __init__(self, state_dict, n_heads, n_margin=32, n_frame=128, n_bin=256, n_note=88, n_velocity=128,
                 gemm_in=None, store=None, dtype=torch.float32, device="cpu"):
        # device: "cpu" (the oracle proper) or a CUDA device for bench.py's "eager PyTorch on the same GPU" baseline leg
        self.device = torch.device(device)
        self.sd = {k: v.detach().to(self.device, dtype) for k, v in state_dict.items()}
        self.h = n_heads
        self.n_margin, self.n_frame, self.n_bin = n_margin, n_frame, n_bin
        self.n_note, self.n_velocity = n_note, n_velocity
        self.hid, self.pf, self.n_enc, self.n_dec, self.ch, self.kw = dims_from_state_dict(state_dict)
        self.q = gemm_in if gemm_in is not None else (lambda t, tag=None: t)
        self.st = store if store is not None else (lambda t: t)
        self.dtype = dtype

    # ---- building blocks -------------------------------------------------------------------
    @staticmethod
    def _t(tag, role):
        """operand tag handed to the gemm_in hook: '<site>:<role>' with role in a/w (linear), q/k/p/v (attention)"""
        return None if tag is None else tag + ':' + role

    def linear(self, x, prefix, tag=None):
        w, b = self.sd[prefix + ".weight"], self.sd[prefix + ".bias"]
        return self.q(x, self._t(tag, 'a')) @ self.q(w, self._t(tag, 'w')).t() + b

    def ln(self, x, prefix):
        return self.st(F.layer_norm(x, (self.hid,), self.sd[prefix + ".weight"], self.sd[prefix + ".bias"], 1e-5))

    def mha(self, q_in, kv_in, prefix, tag=None):
        """model_spec2midi.py:322-360.  q_in [S,Lq,H], kv_in [S,Lk,H] -> ([S,Lq,H], probs [S,h,Lq,Lk])."""
        S, Lq, H = q_in.shape
        d = H // self.h
        Q = self.linear(q_in, prefix + ".fc_q", tag).view(S, Lq, self.h, d).permute(0, 2, 1, 3)
        K = self.linear(kv_in, prefix + ".fc_k", tag).view(S, -1, self.h, d).permute(0, 2, 1, 3)
        V = self.linear(kv_in, prefix + ".fc_v", tag).view(S, -1, self.h, d).permute(0, 2, 1, 3)
        energy = (self.q(Q, self._t(tag, 'q')) @ self.q(K, self._t(tag, 'k')).transpose(-1, -2)) / math.sqrt(d)
        attn = torch.softmax(energy, dim=-1)
        x = self.q(attn, self._t(tag, 'p')) @ self.q(V, self._t(tag, 'v'))
        x = x.permute(0, 2, 1, 3).reshape(S, Lq, H)
        return self.linear(x, prefix + ".fc_o", tag), attn

    def ffn(self, x, prefix, tag=None):
        return self.linear(torch.relu(self.linear(x, prefix + ".fc_1", tag)), prefix + ".fc_2", tag)

    def encoder_layer(self, x, prefix, tag=None):
        a, _ = self.mha(x, x, prefix + ".self_attention", tag)
        x = self.ln(x + a, prefix + ".layer_norm")
        return self.ln(x + self.ffn(x, prefix + ".positionwise_feedforward", tag), prefix + ".layer_norm")

    # ---- encoder ---------------------------------------------------------------------------
    def front(self, spec):
        """model_spec2midi.py:63-95: [B,n_bin,W] -> [B*n_frame, n_bin, H] (before the layer stack)."""
        B = spec.shape[0]
        n_proc = 2 * self.n_margin + 1
        win = spec.unfold(2, n_proc, 1).permute(0, 2, 1, 3)                 # [B,F,bin,65]
        cw = self.sd["encoder_spec2midi.conv.weight"].view(self.ch, self.kw)
        cb = self.sd["encoder_spec2midi.conv.bias"]
        taps = win.unfold(3, self.kw, 1)                                     # [B,F,bin,61,5]
        c = torch.einsum("bfnki,ci->bfnck", taps, cw) + cb.view(1, 1, 1, -1, 1)   # [B,F,bin,4,61]
        c = c.reshape(B * self.n_frame, self.n_bin, self.ch * (n_proc - self.kw + 1))
        e = self.linear(c, "encoder_spec2midi.tok_embedding_freq", "front")
        return e * math.sqrt(self.hid) + self.sd["encoder_spec2midi.pos_embedding_freq.weight"][None]

    def encoder(self, spec):
        x = self.front(spec)
        for i in range(self.n_enc):
            x = self.encoder_layer(x, "encoder_spec2midi.layers_freq.%d" % i, "enc%d" % i)
        return x                                                              # [B*F, bin, H]

    # ---- decoder ---------------------------------------------------------------------------
    def decoder(self, enc, B):
        p = "decoder_spec2midi"
        S = enc.shape[0]
        q0 = self.sd[p + ".pos_embedding_freq.weight"][None].expand(S, -1, -1)
        # DecoderLayer_Zero :255-272
        a, attn = self.mha(q0, enc, p + ".layer_zero_freq.encoder_attention", "dec0")
        t = self.ln(q0 + a, p + ".layer_zero_freq.layer_norm")
        t = self.ln(t + self.ffn(t, p + ".layer_zero_freq.positionwise_feedforward", "dec0"), p + ".layer_zero_freq.layer_norm")
        # DecoderLayer :283-306
        for i in range(self.n_dec - 1):
            lp = p + ".layers_freq.%d" % i
            tag = "dec%d" % (i + 1)
            a, _ = self.mha(t, t, lp + ".self_attention", tag)
            t = self.ln(t + a, lp + ".layer_norm")
            a, attn = self.mha(t, enc, lp + ".encoder_attention", tag)
            t = self.ln(t + a, lp + ".layer_norm")
            t = self.ln(t + self.ffn(t, lp + ".positionwise_feedforward", tag), lp + ".layer_norm")
        attention = attn.reshape(B, self.n_frame, self.h, self.n_note, self.n_bin)
        shp = (B, self.n_frame, self.n_note)
        on_a = torch.sigmoid(self.linear(t, p + ".fc_onset_freq", "headA").reshape(shp))
        off_a = torch.sigmoid(self.linear(t, p + ".fc_offset_freq", "headA").reshape(shp))
        mpe_a = torch.sigmoid(self.linear(t, p + ".fc_mpe_freq", "headA").reshape(shp))
        vel_a = self.linear(t, p + ".fc_velocity_freq", "headA").reshape(shp + (self.n_velocity,))
        # time re-layout :189-191
        u = t.reshape(B, self.n_frame, self.n_note, self.hid).permute(0, 2, 1, 3).reshape(B * self.n_note, self.n_frame, self.hid)
        u = u * math.sqrt(self.hid) + self.sd[p + ".pos_embedding_time.weight"][None]
        for i in range(self.n_dec):
            u = self.encoder_layer(u, p + ".layers_time.%d" % i, "time%d" % i)
        shp_t = (B, self.n_note, self.n_frame)
        on_b = torch.sigmoid(self.linear(u, p + ".fc_onset_time", "headB").reshape(shp_t).permute(0, 2, 1)).contiguous()
        off_b = torch.sigmoid(self.linear(u, p + ".fc_offset_time", "headB").reshape(shp_t).permute(0, 2, 1)).contiguous()
        mpe_b = torch.sigmoid(self.linear(u, p + ".fc_mpe_time", "headB").reshape(shp_t).permute(0, 2, 1)).contiguous()
        vel_b = self.linear(u, p + ".fc_velocity_time", "headB").reshape(shp_t + (self.n_velocity,)).permute(0, 2, 1, 3).contiguous()
        return on_a, off_a, mpe_a, vel_a, attention, on_b, off_b, mpe_b, vel_b

    @torch.no_grad()
    def forward(self, spec):
        """spec [B, n_bin, margin+n_frame+margin] -> the reference's 9-tuple (model_spec2midi.py:35)."""
        spec = torch.as_tensor(spec, dtype=self.dtype).to(self.device)
        return self.decoder(self.encoder(spec), spec.shape[0])

    __call__ = forward


def collapsed_front_weights(sd):
    """The conv + Linear of model_spec2midi.py:73-85 have no non-linearity between them, so they collapse
    to one 65-tap linear map per bin:  e[h] = sum_j Wc[h,j] * win[j] + bc[h]  (SURVEY.md 8a7).
    Returns (Wc [H,65], bc [H]) in fp64-accumulated fp32."""
    cw = sd["encoder_spec2midi.conv.weight"].double()          # [C,1,1,kw]
    cb = sd["encoder_spec2midi.conv.bias"].double()
    W = sd["encoder_spec2midi.tok_embedding_freq.weight"].double()   # [H, C*61]
    b = sd["encoder_spec2midi.tok_embedding_freq.bias"].double()
    C, kw = cw.shape[0], cw.shape[3]
    n_out = W.shape[1] // C
    n_proc = n_out + kw - 1
    H = W.shape[0]
    Wr = W.view(H, C, n_out)
    Wc = torch.zeros(H, n_proc, dtype=torch.float64)
    for i in range(kw):
        Wc[:, i:i + n_out] += torch.einsum("hck,c->hk", Wr, cw[:, 0, 0, i])
    bc = b + torch.einsum("hck,c->h", Wr, cb)
    return Wc.float(), bc.float()


def segment_feature(a_feature, n_frame=128, margin_b=32, margin_f=32, min_value=None):
    """The windowing of AMT.transcript (amt.py:70-73,88-89): pad with min_value, cut [192] windows every
    128 frames.  a_feature [T,n_bin] -> spec [n_seg, n_bin, 192]."""
    import numpy as np
    a = torch.as_tensor(np.asarray(a_feature, dtype=np.float32))
    if min_value is None:
        min_value = float(np.log(np.float32(1e-8)))
    T, nb = a.shape
    n_seg = (T + n_frame - 1) // n_frame
    pad_f = n_seg * n_frame - T + margin_f
    a_in = torch.cat([torch.full((margin_b, nb), min_value), a, torch.full((pad_f, nb), min_value)], 0)
    segs = [a_in[i * n_frame: i * n_frame + margin_b + n_frame + margin_f].t() for i in range(n_seg)]
    return torch.stack(segs, 0).contiguous()

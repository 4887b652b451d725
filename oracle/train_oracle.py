"""CPU restatement of one training step of the reference -- TEST INFRASTRUCTURE (only tests/, smoke() and bench.py's
CPU legs may import this; the product path never does).

  loss      hftt_code/training/train.py:139-151   BCELoss on the six sigmoid outputs + CrossEntropyLoss on the two
                                                  velocity logit tensors, loss = weight_A * loss_A + weight_B * loss_B
  backward  train.py:157                          torch autograd through the forward restatement (oracle/hft_oracle.py)
  Adam      hftt_code/training/m_training.py:146  optim.Adam(model.parameters(), lr) -- restated element-wise below

Pinned against the reference itself: tests/golden/train_reduced.npz (written by oracle/make_golden_train.py from the
unmodified reference modules) holds the loss, every gradient tensor and the parameters after one and two Adam steps.
Dropout is 0 in that fixture (parity runs).
"""
import math

import torch
import torch.nn.functional as F

from . import hft_oracle


def synthetic_labels(B, n_frame=128, n_note=88, n_velocity=128, seed=7):
    """Seeded labels of the shapes train.py:78-80 documents: three float maps in [0,1] (mostly 0/1, some soft), one int64 map."""
    g = torch.Generator().manual_seed(seed)
    u = torch.rand((3, B, n_frame, n_note), generator=g)
    hard = (u > 0.9).float()
    soft = torch.rand((3, B, n_frame, n_note), generator=g)
    lab = torch.where(u > 0.97, soft, hard)
    vel = torch.randint(0, n_velocity, (B, n_frame, n_note), generator=g, dtype=torch.int64)
    vel = torch.where(u[2] > 0.9, vel, torch.zeros_like(vel))
    return lab[0].contiguous(), lab[1].contiguous(), lab[2].contiguous(), vel.contiguous()


def loss_from_outputs(outs, label_onset, label_offset, label_mpe, label_velocity, weight_A=1.0, weight_B=1.0):
    """train.py:117-151 on the 9-tuple."""
    on_a, off_a, mpe_a, vel_a, _, on_b, off_b, mpe_b, vel_b = outs
    yo, yf, ym, yv = label_onset.reshape(-1), label_offset.reshape(-1), label_mpe.reshape(-1), label_velocity.reshape(-1)
    la = F.binary_cross_entropy(on_a.reshape(-1), yo) + F.binary_cross_entropy(off_a.reshape(-1), yf) + \
        F.binary_cross_entropy(mpe_a.reshape(-1), ym) + F.cross_entropy(vel_a.reshape(-1, vel_a.shape[-1]), yv)
    lb = F.binary_cross_entropy(on_b.reshape(-1), yo) + F.binary_cross_entropy(off_b.reshape(-1), yf) + \
        F.binary_cross_entropy(mpe_b.reshape(-1), ym) + F.cross_entropy(vel_b.reshape(-1, vel_b.shape[-1]), yv)
    return weight_A * la + weight_B * lb


class GradOracle(hft_oracle.Oracle):
    """The forward restatement with autograd enabled on every parameter."""

    def __init__(self, state_dict, n_heads, dtype=torch.float32, device="cpu"):
        super().__init__(state_dict, n_heads, dtype=dtype, device=device)
        for v in self.sd.values():
            v.requires_grad_(True)

    def forward_grad(self, spec):
        spec = torch.as_tensor(spec, dtype=self.dtype).to(self.device)
        return self.decoder(self.encoder(spec), spec.shape[0])


def loss_and_grads(state_dict, n_heads, spec, label_onset, label_offset, label_mpe, label_velocity, weight_A=1.0, weight_B=1.0, dtype=torch.float32):
    """-> (loss float, {name: grad tensor}) for one batch."""
    o = GradOracle(state_dict, n_heads, dtype=dtype)
    outs = o.forward_grad(spec)
    loss = loss_from_outputs(outs, label_onset.to(dtype), label_offset.to(dtype), label_mpe.to(dtype), label_velocity, weight_A, weight_B)
    loss.backward()
    return float(loss.detach()), {k: (v.grad.detach().clone() if v.grad is not None else torch.zeros_like(v)) for k, v in o.sd.items()}


def adam_step(params, grads, exp_avg, exp_avg_sq, step, lr=1e-4, beta1=0.9, beta2=0.999, eps=1e-8):
    """torch.optim.Adam's single-tensor update (no weight decay, no amsgrad), in place on dicts of tensors; step counts from 1."""
    bc1 = 1.0 - beta1 ** step
    bc2 = 1.0 - beta2 ** step
    for k in params:
        g = grads[k]
        exp_avg[k].mul_(beta1).add_(g, alpha=1.0 - beta1)
        exp_avg_sq[k].mul_(beta2).addcmul_(g, g, value=1.0 - beta2)
        denom = (exp_avg_sq[k].sqrt() / math.sqrt(bc2)).add_(eps)
        params[k].addcdiv_(exp_avg[k], denom, value=-lr / bc1)


# ---- dropout ------------------------------------------------------------------------------------------------------------
def dropout_multiplier(p, seed, site, n):
    """Restatement of the library's counter-based dropout (csrc/f32_kernels.cuh `drop_keep`): float32 [n] holding 1/(1-p) where
    element idx is kept and 0 where it is dropped; kept iff lowbias32(idx_lo * 0x9E3779B1 ^ idx_hi * 0x85EBCA77 ^ seed ^ site * 0xC2B2AE3D) >= p * 2^32."""
    import numpy as np
    if p <= 0.0:
        return torch.ones(n)
    idx = np.arange(n, dtype=np.uint64)
    m32 = np.uint64(0xFFFFFFFF)
    lo, hi = idx & m32, idx >> np.uint64(32)
    h = ((lo * np.uint64(0x9E3779B1)) & m32) ^ ((hi * np.uint64(0x85EBCA77)) & m32) ^ np.uint64(seed & 0xFFFFFFFF) ^ np.uint64((site * 0xC2B2AE3D) & 0xFFFFFFFF)
    h ^= h >> np.uint64(16)
    h = (h * np.uint64(0x7feb352d)) & m32
    h ^= h >> np.uint64(15)
    h = (h * np.uint64(0x846ca68b)) & m32
    h ^= h >> np.uint64(16)
    thresh = np.uint64(int(float(np.float32(p)) * 4294967296.0))
    keep = h >= thresh
    return torch.from_numpy(np.where(keep, np.float32(1.0) / (np.float32(1.0) - np.float32(p)), np.float32(0.0)).astype(np.float32))


class DropOracle(GradOracle):
    """Training-mode forward of the reference (nn.Dropout active at every site of model_spec2midi.py) with the masks of the
    library's site numbering (include/hft_sm100.h, hft_trainer_set_dropout)."""

    def __init__(self, state_dict, n_heads, p, seed, dtype=torch.float32):
        super().__init__(state_dict, n_heads, dtype=dtype)
        self.p, self.seed = float(p), int(seed)

    def drop(self, x, site):
        m = dropout_multiplier(self.p, self.seed, site, x.numel()).to(x.dtype).view(x.shape)
        return x * m

    def mha_d(self, q_in, kv_in, prefix, site):
        S, Lq, H = q_in.shape
        d = H // self.h
        Q = self.linear(q_in, prefix + ".fc_q").view(S, Lq, self.h, d).permute(0, 2, 1, 3)
        K = self.linear(kv_in, prefix + ".fc_k").view(S, -1, self.h, d).permute(0, 2, 1, 3)
        V = self.linear(kv_in, prefix + ".fc_v").view(S, -1, self.h, d).permute(0, 2, 1, 3)
        attn = torch.softmax((Q @ K.transpose(-1, -2)) / math.sqrt(d), dim=-1).contiguous()
        x = (self.drop(attn, site) @ V).permute(0, 2, 1, 3).reshape(S, Lq, H)
        return self.linear(x, prefix + ".fc_o")

    def ffn_d(self, x, prefix, site):
        return self.linear(self.drop(torch.relu(self.linear(x, prefix + ".fc_1")), site), prefix + ".fc_2")

    def enc_layer_d(self, x, prefix, base):
        x = self.ln(x + self.drop(self.mha_d(x, x, prefix + ".self_attention", base), base + 1), prefix + ".layer_norm")
        return self.ln(x + self.drop(self.ffn_d(x, prefix + ".positionwise_feedforward", base + 2), base + 3), prefix + ".layer_norm")

    def forward_grad(self, spec):
        spec = torch.as_tensor(spec, dtype=self.dtype)
        B = spec.shape[0]
        x = self.drop(self.front(spec).contiguous(), 0)
        for i in range(self.n_enc):
            x = self.enc_layer_d(x, "encoder_spec2midi.layers_freq.%d" % i, 1 + 4 * i)
        enc = x
        p = "decoder_spec2midi"
        S = enc.shape[0]
        q0 = self.sd[p + ".pos_embedding_freq.weight"][None].expand(S, -1, -1)
        d0 = 1 + 4 * self.n_enc
        t = self.ln(q0 + self.drop(self.mha_d(q0, enc, p + ".layer_zero_freq.encoder_attention", d0), d0 + 1), p + ".layer_zero_freq.layer_norm")
        t = self.ln(t + self.drop(self.ffn_d(t, p + ".layer_zero_freq.positionwise_feedforward", d0 + 2), d0 + 3), p + ".layer_zero_freq.layer_norm")
        for i in range(self.n_dec - 1):
            lp, b = p + ".layers_freq.%d" % i, d0 + 4 + 6 * i
            t = self.ln(t + self.drop(self.mha_d(t, t, lp + ".self_attention", b), b + 1), lp + ".layer_norm")
            t = self.ln(t + self.drop(self.mha_d(t, enc, lp + ".encoder_attention", b + 2), b + 3), lp + ".layer_norm")
            t = self.ln(t + self.drop(self.ffn_d(t, lp + ".positionwise_feedforward", b + 4), b + 5), lp + ".layer_norm")
        shp = (B, self.n_frame, self.n_note)
        on_a = torch.sigmoid(self.linear(t, p + ".fc_onset_freq").reshape(shp))
        off_a = torch.sigmoid(self.linear(t, p + ".fc_offset_freq").reshape(shp))
        mpe_a = torch.sigmoid(self.linear(t, p + ".fc_mpe_freq").reshape(shp))
        vel_a = self.linear(t, p + ".fc_velocity_freq").reshape(shp + (self.n_velocity,))
        t0 = d0 + 4 + 6 * (self.n_dec - 1)
        u = t.reshape(B, self.n_frame, self.n_note, self.hid).permute(0, 2, 1, 3).reshape(B * self.n_note, self.n_frame, self.hid)
        u = self.drop((u * math.sqrt(self.hid) + self.sd[p + ".pos_embedding_time.weight"][None]).contiguous(), t0)
        for i in range(self.n_dec):
            u = self.enc_layer_d(u, p + ".layers_time.%d" % i, t0 + 1 + 4 * i)
        shp_t = (B, self.n_note, self.n_frame)
        on_b = torch.sigmoid(self.linear(u, p + ".fc_onset_time").reshape(shp_t).permute(0, 2, 1)).contiguous()
        off_b = torch.sigmoid(self.linear(u, p + ".fc_offset_time").reshape(shp_t).permute(0, 2, 1)).contiguous()
        mpe_b = torch.sigmoid(self.linear(u, p + ".fc_mpe_time").reshape(shp_t).permute(0, 2, 1)).contiguous()
        vel_b = self.linear(u, p + ".fc_velocity_time").reshape(shp_t + (self.n_velocity,)).permute(0, 2, 1, 3).contiguous()
        return on_a, off_a, mpe_a, vel_a, None, on_b, off_b, mpe_b, vel_b


def loss_and_grads_dropout(state_dict, n_heads, spec, label_onset, label_offset, label_mpe, label_velocity, p, seed, weight_A=1.0, weight_B=1.0,
                           dtype=torch.float32):
    o = DropOracle(state_dict, n_heads, p, seed, dtype=dtype)
    outs = o.forward_grad(spec)
    loss = loss_from_outputs(outs, label_onset.to(dtype), label_offset.to(dtype), label_mpe.to(dtype), label_velocity, weight_A, weight_B)
    loss.backward()
    return float(loss.detach()), {k: (v.grad.detach().clone() if v.grad is not None else torch.zeros_like(v)) for k, v in o.sd.items()}

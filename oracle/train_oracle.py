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

    def __init__(self, state_dict, n_heads, dtype=torch.float32):
        super().__init__(state_dict, n_heads, dtype=dtype)
        for v in self.sd.values():
            v.requires_grad_(True)

    def forward_grad(self, spec):
        spec = torch.as_tensor(spec, dtype=self.dtype)
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

"""Writes tests/golden/train_reduced.npz from the UNMODIFIED reference modules (build container only; /root/reference is
absent on the GPU box).  One training iteration of hftt_code/training/train.py:89-158 on the reduced model with the
weights of tests/golden/hft_reduced.npz, B = 2 segments of that fixture, seeded labels, dropout 0, Adam(lr 1e-4):
loss, every gradient, parameters after step 1, and the loss of a second iteration.

    python -m oracle.make_golden_train
"""
import os
import sys

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import _refload, train_oracle  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def main():
    ref_amt, ref_model = _refload.load()
    cfg = _refload.config()
    g = np.load(os.path.join(OUT, "hft_reduced.npz"))
    model = _refload.build_model(ref_model, cfg, 64, 128, 2, 2)
    sd = {k[2:]: torch.from_numpy(g[k]).clone() for k in g.files if k.startswith("w:")}
    model.load_state_dict(sd)
    for m in model.modules():
        if isinstance(m, nn.Dropout):
            m.p = 0.0
    model.train()
    spec = torch.from_numpy(g["spec"][:2]).clone()
    yo, yf, ym, yv = train_oracle.synthetic_labels(2)
    bce = [nn.BCELoss() for _ in range(6)]
    ce = [nn.CrossEntropyLoss() for _ in range(2)]
    opt = torch.optim.Adam(model.parameters(), lr=1e-4)
    d = {"spec": spec.numpy(), "label_onset": yo.numpy(), "label_offset": yf.numpy(), "label_mpe": ym.numpy(), "label_velocity": yv.numpy(),
         "lr": np.float32(1e-4), "weight_A": np.float32(1.0), "weight_B": np.float32(1.0)}

    def iteration():                       # train.py:104-158
        opt.zero_grad()
        on_a, off_a, mpe_a, vel_a, _, on_b, off_b, mpe_b, vel_b = model(spec)
        la = bce[0](on_a.contiguous().view(-1), yo.view(-1)) + bce[1](off_a.contiguous().view(-1), yf.view(-1)) + \
            bce[2](mpe_a.contiguous().view(-1), ym.view(-1)) + ce[0](vel_a.contiguous().view(-1, vel_a.shape[-1]), yv.view(-1))
        lb = bce[3](on_b.contiguous().view(-1), yo.view(-1)) + bce[4](off_b.contiguous().view(-1), yf.view(-1)) + \
            bce[5](mpe_b.contiguous().view(-1), ym.view(-1)) + ce[1](vel_b.contiguous().view(-1, vel_b.shape[-1]), yv.view(-1))
        loss = 1.0 * la + 1.0 * lb
        loss.backward()
        return loss

    loss = iteration()
    d["loss"] = np.float32(loss.item())
    for k, p in model.named_parameters():
        d["g:" + k] = p.grad.detach().numpy().copy()
    opt.step()
    for k, p in model.named_parameters():
        d["p1:" + k] = p.detach().numpy().copy()
    loss2 = iteration()
    d["loss2"] = np.float32(loss2.item())
    opt.step()
    d["p2_checksum"] = np.array([float(p.detach().double().sum()) for _, p in model.named_parameters()])
    np.savez_compressed(os.path.join(OUT, "train_reduced.npz"), **d)
    print("loss %.6f loss2 %.6f, %d gradient tensors" % (loss.item(), loss2.item(), sum(1 for k in d if k.startswith("g:"))))


if __name__ == "__main__":
    main()

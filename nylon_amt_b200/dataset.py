"""Training data on the device -- mirror of the reference's hftt_code/training/dataset.py (class MyDataset) for the B200 training step.

Same constructor arguments and the same on-disk format (the pickles hftt_code/corpus/make_dataset.py:85-160 writes: one float32 feature
matrix [T_total, 256] per split, label matrices [T_total, 88], and the array of window start frames).  `__getitem__` returns exactly
what the reference returns (CPU tensors), so a torch DataLoader keeps working; the B200 path is `to_device()` + `batches()`:
the matrices are uploaded ONCE and every batch is gathered on the device (no per-sample Python slicing, no worker processes, no H2D per
step), as the [B, 256, 192] strided view hft_train_forward_backward takes.
"""
import pickle

import numpy as np
import torch


def _load(path):
    with open(path, 'rb') as f:
        return pickle.load(f)


class MyDataset(torch.utils.data.Dataset):
    def __init__(self, f_feature, f_label_onset, f_label_offset, f_label_mpe, f_label_velocity, f_idx, config, n_slice):
        super().__init__()
        self.feature = torch.from_numpy(np.asarray(_load(f_feature)))
        self.label_onset = torch.from_numpy(np.asarray(_load(f_label_onset)))
        self.label_offset = torch.from_numpy(np.asarray(_load(f_label_offset)))
        self.label_mpe = torch.from_numpy(np.asarray(_load(f_label_mpe)))
        self.flag_velocity = f_label_velocity is not None
        if self.flag_velocity:
            self.label_velocity = torch.from_numpy(np.asarray(_load(f_label_velocity)))
        idx = torch.from_numpy(np.asarray(_load(f_idx)))
        if n_slice > 1:                                      # dataset.py:35-38: keep every n_slice-th window
            idx = idx[:int(len(idx) / n_slice) * n_slice][::n_slice]
        self.idx = idx
        self.config = config
        self.data_size = len(self.idx)
        self._dev = None

    def __len__(self):
        return self.data_size

    def __getitem__(self, idx):                              # dataset.py:46-74, unchanged semantics (CPU tensors)
        c = self.config['input']
        s = int(self.idx[idx])
        spec = self.feature[s - c['margin_b']: s + c['num_frame'] + c['margin_f']].T
        lab = slice(s, s + c['num_frame'])
        out = (spec, self.label_onset[lab], self.label_offset[lab], self.label_mpe[lab].float())
        return out + (self.label_velocity[lab].long(),) if self.flag_velocity else out

    # ---- device path -----------------------------------------------------------------------------------------------------
    def to_device(self, device='cuda'):
        """Upload the split once; labels are converted to the dtypes the training step reads (float32 / int64)."""
        dev = torch.device(device)
        if dev.type != 'cuda':
            raise RuntimeError("the B200 data path is CUDA only (no CPU fallback)")
        self._dev = {
            'feature': self.feature.to(dev, torch.float32).contiguous(),
            'onset': self.label_onset.to(dev, torch.float32).contiguous(),
            'offset': self.label_offset.to(dev, torch.float32).contiguous(),
            'mpe': self.label_mpe.to(dev, torch.float32).contiguous(),
            'velocity': self.label_velocity.to(dev, torch.int64).contiguous() if self.flag_velocity else None,
            'idx': self.idx.to(dev, torch.int64),
        }
        c = self.config['input']
        self._win = torch.arange(-c['margin_b'], c['num_frame'] + c['margin_f'], device=dev)
        self._lab = torch.arange(0, c['num_frame'], device=dev)
        return self

    def gather(self, sample_indices):
        """Batch of windows for the given dataset indices (device LongTensor): spec [B, 256, 192] (a strided view, time-contiguous
        rows like the reference's `.T`), label_onset / offset / mpe [B, 128, 88] float32, label_velocity [B, 128, 88] int64."""
        if self._dev is None:
            raise RuntimeError("call to_device() first")
        d = self._dev
        start = d['idx'][sample_indices]
        rows = start[:, None] + self._win[None]                           # [B, 192] feature rows
        spec = d['feature'][rows].transpose(1, 2)                         # [B, 192, 256] gathered, viewed as [B, 256, 192]
        lrows = start[:, None] + self._lab[None]
        out = (spec, d['onset'][lrows], d['offset'][lrows], d['mpe'][lrows])
        return out + (d['velocity'][lrows],) if self.flag_velocity else out

    def batches(self, batch_size, shuffle=True, drop_last=True, generator=None, rank=0, world=1):
        """Iterator of device batches (what DataLoader(dataset, batch_size, shuffle) yields, m_training.py:96-104).  Under data
        parallelism every rank draws the same permutation (same generator seed) and takes its contiguous share of each global batch."""
        if self._dev is None:
            self.to_device()
        dev = self._dev['idx'].device
        n = self.data_size
        order = torch.randperm(n, generator=generator) if shuffle else torch.arange(n)
        order = order.to(dev)
        step = batch_size * world
        stop = n - (n % step) if drop_last else n
        for b0 in range(0, stop, step):
            sel = order[b0 + rank * batch_size: b0 + (rank + 1) * batch_size]
            if sel.numel() == 0:
                break
            yield self.gather(sel)

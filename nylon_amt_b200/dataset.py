"""Training data on the device -- mirror of the reference's hftt_code/training/dataset.py (class MyDataset) for the B200 training step.

Same constructor arguments and the same on-disk format (the pickles hftt_code/corpus/make_dataset.py:85-160 writes: one float32 feature
matrix [T_total, 256] per split, label matrices [T_total, 88], and the array of window start frames).  `__getitem__` returns exactly
what the reference returns (CPU tensors), so a torch DataLoader keeps working; the B200 path is `to_device()` + `batches()`:
the matrices are uploaded ONCE and every batch is gathered on the device (no per-sample Python slicing, no worker processes, no H2D per
step), as the [B, 256, 192] strided view hft_train_forward_backward takes.

Raw layout (B200 path): a MAESTRO-size split is tens of GB; unpickling it builds the whole matrix on the host heap before a single byte
moves.  `convert_to_raw()` rewrites every pickle of a split ONCE as a flat `.npy` (same dtype, same shape, C order -- the reference's
arrays byte for byte behind a 128-byte header); MyDataset accepts those paths and memory-maps them, so construction is O(1),
`__getitem__` pages in only the 192 rows it touches, and `to_device()` streams the file to HBM through one pinned staging buffer
(page cache -> pinned -> device) without ever holding a host copy.
"""
import os
import pickle

import numpy as np
import torch

RAW_EXT = '.npy'
_STAGE_BYTES = 64 << 20            # pinned staging buffer of to_device()


def _load(path):
    """A split array: memory-mapped when it is a raw `.npy` (convert_to_raw), unpickled otherwise (make_dataset.py's own format)."""
    if str(path).endswith(RAW_EXT):
        return np.load(path, mmap_mode='r')
    with open(path, 'rb') as f:
        return pickle.load(f)


def convert_to_raw(f_pickle, f_raw=None):
    """make_dataset.py:85-160 pickle -> flat `.npy` next to it (or at f_raw); returns the raw path.  dtype / shape / values unchanged."""
    f_raw = f_raw or (os.path.splitext(str(f_pickle))[0] + RAW_EXT)
    with open(f_pickle, 'rb') as f:
        a = np.ascontiguousarray(pickle.load(f))
    tmp = f_raw + '.tmp'
    with open(tmp, 'wb') as f:
        np.save(f, a, allow_pickle=False)
    os.replace(tmp, f_raw)
    return f_raw


def _as_tensor(a):
    """torch view of a split array; a read-only memory map is wrapped without copying (never written through)."""
    if isinstance(a, np.memmap):
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")              # "the given NumPy array is not writable": the dataset only reads
            return torch.from_numpy(a)
    return torch.from_numpy(np.asarray(a))


def _upload(t, dev, dtype):
    """CPU tensor (possibly backed by a memory map) -> device tensor of `dtype`, streamed in _STAGE_BYTES pieces through pinned memory."""
    out = torch.empty(t.shape, device=dev, dtype=dtype)
    if t.numel() == 0:
        return out
    rows = max(1, _STAGE_BYTES // max(1, t[0].numel() * t.element_size()))
    stage = [torch.empty((min(rows, t.shape[0]),) + tuple(t.shape[1:]), dtype=t.dtype).pin_memory() for _ in range(2)]
    events = [None, None]
    for k, r0 in enumerate(range(0, t.shape[0], rows)):
        r1 = min(t.shape[0], r0 + rows)
        buf = stage[k & 1]
        if events[k & 1] is not None:
            events[k & 1].synchronize()                  # the copy that last used this buffer has left it
        buf[:r1 - r0].copy_(t[r0:r1])                    # page cache -> pinned
        out[r0:r1].copy_(buf[:r1 - r0], non_blocking=True)      # pinned -> HBM (+ dtype conversion on the device)
        events[k & 1] = torch.cuda.Event()
        events[k & 1].record()
    torch.cuda.current_stream(dev).synchronize()
    return out


class MyDataset(torch.utils.data.Dataset):
    def __init__(self, f_feature, f_label_onset, f_label_offset, f_label_mpe, f_label_velocity, f_idx, config, n_slice):
        super().__init__()
        self.feature = _as_tensor(_load(f_feature))
        self.label_onset = _as_tensor(_load(f_label_onset))
        self.label_offset = _as_tensor(_load(f_label_offset))
        self.label_mpe = _as_tensor(_load(f_label_mpe))
        self.flag_velocity = f_label_velocity is not None
        if self.flag_velocity:
            self.label_velocity = _as_tensor(_load(f_label_velocity))
        idx = torch.from_numpy(np.array(_load(f_idx)))       # small: always copied into memory
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
        with torch.cuda.device(dev):
            self._dev = {
                'feature': _upload(self.feature, dev, torch.float32),
                'onset': _upload(self.label_onset, dev, torch.float32),
                'offset': _upload(self.label_offset, dev, torch.float32),
                'mpe': _upload(self.label_mpe, dev, torch.float32),
                'velocity': _upload(self.label_velocity, dev, torch.int64) if self.flag_velocity else None,
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
        if shuffle and generator is None and world > 1:
            # every rank must draw the SAME permutation: without a shared generator the ranks' global RNGs would hand them overlapping
            # samples.  Derive one from the epoch counter (identical on all ranks, advancing every call).
            self._epoch = getattr(self, '_epoch', 0) + 1
            generator = torch.Generator().manual_seed(0x5EED0000 + self._epoch)
        order = torch.randperm(n, generator=generator) if shuffle else torch.arange(n)
        order = order.to(dev)
        step = batch_size * world
        stop = n - (n % step) if drop_last else n
        for b0 in range(0, stop, step):
            sel = order[b0 + rank * batch_size: b0 + (rank + 1) * batch_size]
            if sel.numel() == 0:
                break
            yield self.gather(sel)

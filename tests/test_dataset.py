"""hftt_code/training/dataset.py mirror: same items as the reference class on the same pickles (CPU), device gather equals them (GPU)."""
import os
import pickle

import numpy as np
import pytest
import torch

import nylon_amt_b200 as hft
from nylon_amt_b200.dataset import MyDataset


@pytest.fixture(scope="module")
def pickles(tmp_path_factory):
    d = tmp_path_factory.mktemp("ds")
    rng = np.random.default_rng(3)
    T = 2000
    files = {}
    arrs = {"feature": rng.standard_normal((T, 256)).astype(np.float32), "onset": rng.random((T, 88)).astype(np.float32),
            "offset": rng.random((T, 88)).astype(np.float32), "mpe": rng.random((T, 88)) > 0.7,
            "velocity": rng.integers(0, 128, (T, 88)).astype(np.int8), "idx": np.arange(32, T - 200, 7).astype(np.int32)}
    for k, a in arrs.items():
        files[k] = str(d / (k + ".pkl"))
        with open(files[k], "wb") as f:
            pickle.dump(a, f, protocol=4)
    return files, arrs


def _make(files, n_slice=1):
    return MyDataset(files["feature"], files["onset"], files["offset"], files["mpe"], files["velocity"], files["idx"], hft.default_config(), n_slice)


def test_items_follow_the_reference_slicing(pickles):
    files, a = pickles
    ds = _make(files)
    assert len(ds) == len(a["idx"])
    for k in (0, 5, len(ds) - 1):
        s = int(a["idx"][k])
        spec, on, off, mpe, vel = ds[k]
        assert tuple(spec.shape) == (256, 192) and not spec.is_contiguous()          # the reference hands a .T view to the model (dataset.py:56)
        assert np.array_equal(spec.numpy(), a["feature"][s - 32:s + 160].T)
        assert np.array_equal(on.numpy(), a["onset"][s:s + 128]) and np.array_equal(off.numpy(), a["offset"][s:s + 128])
        assert mpe.dtype == torch.float32 and np.array_equal(mpe.numpy(), a["mpe"][s:s + 128].astype(np.float32))
        assert vel.dtype == torch.int64 and np.array_equal(vel.numpy(), a["velocity"][s:s + 128].astype(np.int64))
    ds3 = _make(files, n_slice=3)
    n = len(a["idx"])
    assert np.array_equal(ds3.idx.numpy(), a["idx"][:n // 3 * 3][::3])


def test_raw_memory_mapped_layout_equals_pickles(pickles):
    """convert_to_raw: every split array as a flat .npy; MyDataset memory-maps it (no host copy) and returns the same items."""
    from nylon_amt_b200.dataset import convert_to_raw
    files, a = pickles
    raw = {k: convert_to_raw(v) for k, v in files.items()}
    for k, path in raw.items():
        m = np.load(path, mmap_mode="r")
        assert isinstance(m, np.memmap) and m.dtype == a[k].dtype and m.shape == a[k].shape and np.array_equal(m, a[k])
        assert os.path.getsize(path) == a[k].nbytes + 128                         # the reference's bytes behind one header
    ds_raw, ds_pkl = _make(raw), _make(files)
    assert isinstance(ds_raw.feature.numpy().base, np.memmap) or ds_raw.feature.data_ptr() != ds_pkl.feature.data_ptr()
    assert len(ds_raw) == len(ds_pkl)
    for k in (0, 9, len(ds_raw) - 1):
        for x, y in zip(ds_raw[k], ds_pkl[k]):
            assert x.dtype == y.dtype and x.shape == y.shape and torch.equal(x, y)
    assert torch.equal(_make(raw, n_slice=3).idx, _make(files, n_slice=3).idx)


def test_reference_class_agrees_when_available(pickles):
    from oracle import _refload
    if not _refload.available():
        pytest.skip("reference tree not present")
    import sys
    _refload.load()
    sys.path.insert(0, os.path.join(_refload.REF_CODE, "training"))
    import dataset as ref_dataset
    files, _ = pickles
    cfg = hft.default_config()
    ref = ref_dataset.MyDataset(files["feature"], files["onset"], files["offset"], files["mpe"], files["velocity"], files["idx"], cfg, 2)
    mine = _make(files, n_slice=2)
    assert len(ref) == len(mine)
    for k in (0, 3, len(ref) - 1):
        for x, y in zip(ref[k], mine[k]):
            assert x.dtype == y.dtype and torch.equal(x, y)


@pytest.mark.gpu
def test_device_batches_equal_items(pickles):
    files, _ = pickles
    ds = _make(files).to_device()
    sel = torch.tensor([0, 17, 3, len(ds) - 1], device="cuda")
    batch = ds.gather(sel)
    assert tuple(batch[0].shape) == (4, 256, 192) and batch[0].stride(2) == 256          # strided view, no transpose copy
    for b, k in enumerate(sel.tolist()):
        for got, ref in zip(batch, ds[k]):
            assert torch.equal(got[b].cpu(), ref)
    g = torch.Generator().manual_seed(1)
    seen = [b[0].shape[0] for b in ds.batches(8, shuffle=True, generator=g)]
    assert all(s == 8 for s in seen) and len(seen) == len(ds) // 8
    # data parallel: two ranks with the same seed split every global batch without overlap
    r0 = [b[1] for b in ds.batches(4, generator=torch.Generator().manual_seed(2), rank=0, world=2)]
    r1 = [b[1] for b in ds.batches(4, generator=torch.Generator().manual_seed(2), rank=1, world=2)]
    full = [b[1] for b in ds.batches(8, generator=torch.Generator().manual_seed(2))]
    assert len(r0) == len(r1) == len(full) and torch.equal(torch.cat([r0[0], r1[0]]), full[0])


@pytest.mark.gpu
def test_raw_layout_streams_to_the_device(pickles):
    """to_device() from the memory-mapped files (staged through pinned memory in pieces) equals the upload of the unpickled arrays."""
    from nylon_amt_b200 import dataset as dsmod
    files, _ = pickles
    raw = {k: dsmod.convert_to_raw(v) for k, v in files.items()}
    old = dsmod._STAGE_BYTES
    dsmod._STAGE_BYTES = 300 * 1024                  # force many pieces (2000 x 256 floats = 2 MB)
    try:
        a = _make(raw).to_device()
    finally:
        dsmod._STAGE_BYTES = old
    b = _make(files).to_device()
    for k in ("feature", "onset", "offset", "mpe", "velocity", "idx"):
        assert a._dev[k].dtype == b._dev[k].dtype and torch.equal(a._dev[k], b._dev[k]), k
    sel = torch.tensor([1, 40, 7], device="cuda")
    for x, y in zip(a.gather(sel), b.gather(sel)):
        assert torch.equal(x, y)


@pytest.mark.gpu
def test_training_step_consumes_device_batches(pickles, golden_dir):
    files, _ = pickles
    ds = _make(files).to_device()
    g = np.load(os.path.join(golden_dir, "hft_reduced.npz"))
    model = hft.build_model(hft.default_config(), 64, 128, 2, 2, dropout=0.0, device="cuda")
    model.load_state_dict({k[2:]: torch.from_numpy(g[k]).clone() for k in g.files if k.startswith("w:")})
    opt = hft.training.Adam(model, lr=1e-3, batch_size=2)
    losses = []
    for i, batch in enumerate(ds.batches(2, generator=torch.Generator().manual_seed(0))):
        losses.append(float(hft.training.train_step(model, opt, *batch).item()))
        if i == 5:
            break
    assert all(np.isfinite(losses)) and losses[-1] < losses[0]

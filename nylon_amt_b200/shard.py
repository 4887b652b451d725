"""Multi-GPU sharding of the hot path: files / clips / segments are independent, so every rank takes a
contiguous block and there is NO data-path collective -- the only communication is the reduction of a few scalars
(audio seconds, max device time) for the report.  One process per GPU (torchrun); works with any torch.distributed
backend (nccl on GPUs, gloo in the CPU tests)."""
import os


def world():
    """(rank, local_rank, world_size) from the torchrun environment (1 process when not launched by torchrun)."""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def partition(n_items, world_size, rank):
    """Contiguous block [lo, hi) of rank `rank`: sizes differ by at most one, blocks cover 0..n_items exactly once."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank %d / world %d" % (rank, world_size))
    base, extra = divmod(n_items, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def reduce_report(local_audio_s, local_ms, device=None):
    """Whole-job numbers for the report: sum of audio seconds and MAX over ranks of the device time."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(local_audio_s), float(local_ms)
    a = torch.tensor([float(local_audio_s)], dtype=torch.float64, device=device)
    t = torch.tensor([float(local_ms)], dtype=torch.float64, device=device)
    dist.all_reduce(a, op=dist.ReduceOp.SUM)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(a.item()), float(t.item())


def features_sharded(amt, wav_paths):
    """The file loop of hftt_code/corpus/conv_wav2fe.py:41-48 split across ranks: returns {path: feature} for this
    rank's block only (each rank pickles its own outputs, like the reference does per file)."""
    rank, _, ws = world()
    lo, hi = partition(len(wav_paths), ws, rank)
    return {p: amt.wav2feature(p) for p in wav_paths[lo:hi]}


# ---- data-parallel training (BASELINE config 5): the one exchange step of the whole path -------------------------------
def block_range(n_items, rank, world_size):
    """Alias of partition() with the (rank, world) argument order the training code reads naturally."""
    return partition(n_items, world_size, rank)


def flatten_bucket(tensors):
    """Concatenate gradient tensors into ONE flat fp32 bucket (the CUDA trainer's gradients already are one)."""
    import torch
    return torch.cat([t.reshape(-1).to(torch.float32) for t in tensors])


def unflatten_bucket(flat, like):
    """Views of `flat` shaped like the tensors of `like`."""
    out, off = [], 0
    for t in like:
        out.append(flat[off:off + t.numel()].view(t.shape))
        off += t.numel()
    return out


def allreduce_bucket(flat, group=None):
    """Sum the flat gradient bucket over the ranks in place (NCCL over NVLink on GPUs, gloo in the CPU tests); returns the
    world size to divide by.  A single process is a no-op."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        return dist.get_world_size(group)
    return 1

"""Multi-GPU sharding of the hot path: one process per GPU, independent units (frame pairs for KLT, reference-descriptor
rows for matching) split into contiguous blocks, NO collective on the data path; only the final per-feature results are
gathered.  The partition / gather logic is backend agnostic: `gloo` on CPU tensors in the tests,
`nccl` (over NVLink) or plain host gathers in production.
"""
import numpy as np


def shard_bounds(n_units, world_size, rank):
    """Contiguous block of units owned by `rank`: unit u belongs to rank u * world_size // n_units (balanced to +-1)."""
    lo = (n_units * rank + world_size - 1) // world_size
    hi = (n_units * (rank + 1) + world_size - 1) // world_size
    return lo, hi


def shard_feature_batch(feat_offsets, world_size, rank):
    """Splits a batch of frame pairs.  Returns (pair_lo, pair_hi, feat_lo, feat_hi, local_offsets)."""
    feat_offsets = np.asarray(feat_offsets, dtype=np.int64)
    n_pairs = feat_offsets.shape[0] - 1
    lo, hi = shard_bounds(n_pairs, world_size, rank)
    f_lo, f_hi = int(feat_offsets[lo]), int(feat_offsets[hi])
    return lo, hi, f_lo, f_hi, (feat_offsets[lo:hi + 1] - f_lo).astype(np.int32)


def gather_to_rank0(local_arrays, group=None):
    """Gathers a tuple of per-rank numpy arrays (concatenated along axis 0 in rank order) on rank 0; other ranks get None.
    Uses torch.distributed when a process group is initialised, otherwise returns the local arrays (single process)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return tuple(np.asarray(a) for a in local_arrays)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    backend = dist.get_backend(group)
    device = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    out = []
    for a in local_arrays:
        a = np.ascontiguousarray(a)
        # variable block sizes: exchange the row counts first, pad to the maximum
        n_local = torch.tensor([a.shape[0]], dtype=torch.int64, device=device)
        counts = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
        dist.all_gather(counts, n_local, group=group)
        counts = [int(c.item()) for c in counts]
        n_max = max(counts)
        row_shape = a.shape[1:]
        t = torch.zeros((n_max,) + row_shape, dtype=torch.from_numpy(a[:0]).dtype, device=device)
        if a.shape[0]:
            t[:a.shape[0]] = torch.from_numpy(a).to(device)
        if rank == 0:
            bufs = [torch.zeros_like(t) for _ in range(world)]
            dist.gather(t, gather_list=bufs, dst=0, group=group)
            out.append(np.concatenate([b[:c].cpu().numpy() for b, c in zip(bufs, counts)], axis=0))
        else:
            dist.gather(t, gather_list=None, dst=0, group=group)
            out.append(None)
    return tuple(out)


def track_sharded(track_fn, feat_offsets, ref_uv, world_size, rank, group=None):
    """Runs `track_fn(pair_lo, pair_hi, local_offsets, local_ref_uv) -> (cur_uv, status)` on this rank's block of frame
    pairs and gathers (cur_uv, status) in global feature order on rank 0."""
    lo, hi, f_lo, f_hi, local_offsets = shard_feature_batch(feat_offsets, world_size, rank)
    ref_uv = np.asarray(ref_uv, dtype=np.float32).reshape(-1, 2)
    if hi > lo and f_hi > f_lo:
        cur_uv, status = track_fn(lo, hi, local_offsets, ref_uv[f_lo:f_hi])
    else:
        cur_uv, status = np.zeros((0, 2), np.float32), np.zeros((0,), np.uint8)
    return gather_to_rank0((np.asarray(cur_uv, np.float32), np.asarray(status, np.uint8)), group)


def match_sharded(match_fn, n_ref, world_size, rank, group=None):
    """Runs `match_fn(row_lo, row_hi) -> idx[row_hi - row_lo]` on this rank's block of reference rows (the current set is
    replicated on every GPU) and gathers the index vector on rank 0."""
    lo, hi = shard_bounds(n_ref, world_size, rank)
    idx = match_fn(lo, hi) if hi > lo else np.zeros((0,), np.int32)
    return gather_to_rank0((np.asarray(idx, np.int32),), group)[0]

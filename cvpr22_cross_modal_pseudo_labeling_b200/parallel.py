"""Image sharding across the GPUs of one box (one process per GPU, torch.distributed).

Every stage of the RoI hot path is independent per image, so ranks take contiguous blocks of
images and never exchange features (the reference shards the same way through
DistributedSampler, data/build.py:66-68).  The only collective is an all-gather of the
fixed-size pseudo-label records -- the counterpart of the reference's pickled all_gather of
predictions (utils/comm.py:48-88).
"""
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n_items, rank=None, world_size=None):
    """Contiguous block [lo, hi) of `n_items` images owned by `rank` (remainder to the first ranks)."""
    if rank is None or world_size is None:
        rank, world_size = world()
    base, rem = divmod(n_items, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def all_gather_records(records, counts, sizes=None):
    """records [B_local, W, 8] fp32, counts [B_local] int32 -> the same for ALL images, rank-major.
    Shards may differ in size by one image; shorter shards are padded and trimmed after.
    `sizes` (images per rank) skips the size exchange and its host sync when the caller knows
    the sharding (shard_range is deterministic)."""
    rank, ws = world()
    if ws == 1:
        return records, counts
    if sizes is None:
        n_local = torch.tensor([records.shape[0]], device=records.device, dtype=torch.int64)
        sz = [torch.zeros_like(n_local) for _ in range(ws)]
        dist.all_gather(sz, n_local)
        sizes = [int(s.item()) for s in sz]
    if len(set(sizes)) == 1:
        # equal shards: records and counts travel in ONE collective (counts ride along as fp32,
        # exact for any realistic count), no padding, no host sync
        nrec = records.numel()
        flat = torch.cat([records.reshape(-1), counts.to(records.dtype)])
        out = torch.empty((ws * flat.numel(),), dtype=records.dtype, device=records.device)
        dist.all_gather_into_tensor(out, flat)
        out = out.view(ws, -1)
        out_r = out[:, :nrec].reshape((ws * sizes[0],) + tuple(records.shape[1:]))
        out_c = out[:, nrec:].reshape(-1).to(counts.dtype)
        return out_r, out_c
    n_max = max(sizes)
    pad_r = torch.zeros((n_max,) + tuple(records.shape[1:]), dtype=records.dtype, device=records.device)
    pad_c = torch.zeros((n_max,), dtype=counts.dtype, device=counts.device)
    pad_r[: records.shape[0]] = records
    pad_c[: counts.shape[0]] = counts
    out_r = torch.empty((ws * n_max,) + tuple(records.shape[1:]), dtype=records.dtype, device=records.device)
    out_c = torch.empty((ws * n_max,), dtype=counts.dtype, device=counts.device)
    dist.all_gather_into_tensor(out_r, pad_r)
    dist.all_gather_into_tensor(out_c, pad_c)
    keep = torch.cat([torch.arange(r * n_max, r * n_max + s, device=records.device) for r, s in enumerate(sizes)])
    return out_r[keep], out_c[keep]

"""Multi-GPU plumbing for batches that shard by alignment.

Alignments are independent, so the data path needs no collective: the read batch is cut into
contiguous per-rank ranges, the shared target is broadcast once from rank 0 and the fixed-size
per-read results are gathered on rank 0.  torch.distributed is used for exactly these two
transfers (NCCL on GPUs; the same code runs on gloo/CPU tensors, which is how the CPU tests cover
the N>1 path)."""
import numpy as np
import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """Contiguous [lo, hi) of item indices owned by `rank` (sizes differ by at most one)."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def broadcast_target(target, length, device, src=0):
    """Rank `src` passes its uint8 numpy target; every rank gets a host copy (one broadcast)."""
    rank = dist.get_rank()
    buf = torch.from_numpy(target).to(device) if rank == src else torch.empty(length, dtype=torch.uint8, device=device)
    dist.broadcast(buf, src=src)
    return buf.cpu().numpy()


def gather_int32(values, device, dst=0):
    """Gathers one int32 numpy vector per rank on `dst` (ragged lengths allowed); returns the
    concatenation in rank order on `dst`, None elsewhere."""
    rank, world = dist.get_rank(), dist.get_world_size()
    mine = torch.from_numpy(np.ascontiguousarray(values, dtype=np.int32)).to(device)
    sizes = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([mine.numel()], dtype=torch.int64, device=device))
    cap = int(max(int(s.item()) for s in sizes))
    padded = torch.zeros(cap, dtype=torch.int32, device=device)
    padded[:mine.numel()] = mine
    bufs = [torch.empty(cap, dtype=torch.int32, device=device) for _ in range(world)] if rank == dst else None
    dist.gather(padded, bufs, dst=dst)
    if rank != dst:
        return None
    return np.concatenate([b[:int(s.item())].cpu().numpy() for b, s in zip(bufs, sizes)])


def broadcast_target_into(target, out, device, src=0):
    """As broadcast_target, but every rank receives the bytes in its own host buffer `out` (uint8 numpy, e.g. pinned):
    the target keeps ONE address per rank from step to step, so pointer arrays built on it stay valid."""
    rank = dist.get_rank()
    buf = torch.from_numpy(target).to(device) if rank == src else torch.empty(len(out), dtype=torch.uint8, device=device)
    dist.broadcast(buf, src=src)
    torch.from_numpy(out).copy_(buf)
    return out


def gather_int32_known(values, counts, device, out=None, dst=0):
    """gather_int32 when every rank's count is known up front (shard_range): one gather, no size exchange; `out`
    (int32 numpy of sum(counts), e.g. pinned) receives the concatenation on `dst`."""
    rank, world = dist.get_rank(), dist.get_world_size()
    cap = max(counts)
    padded = torch.zeros(cap, dtype=torch.int32, device=device)
    padded[:len(values)] = torch.from_numpy(np.ascontiguousarray(values, dtype=np.int32)).to(device, non_blocking=True)
    bufs = [torch.empty(cap, dtype=torch.int32, device=device) for _ in range(world)] if rank == dst else None
    dist.gather(padded, bufs, dst=dst)
    if rank != dst:
        return None
    if out is None:
        out = np.empty(sum(counts), dtype=np.int32)
    at = 0
    for b, c in zip(bufs, counts):
        torch.from_numpy(out[at:at + c]).copy_(b[:c])
        at += c
    return out

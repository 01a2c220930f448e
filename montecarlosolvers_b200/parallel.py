"""Multi-GPU sharding of independent anneals (one process per GPU, torch.distributed).

Anneals / restarts / reads never interact: the replica axis is split contiguously
across ranks, the compiled instance is replicated, and the Philox counter carries the GLOBAL replica
index, so results do not depend on the number of GPUs.  The only communication is one all-gather of
the per-replica final energies and a broadcast of the best configuration from its owner (NCCL on
GPUs, gloo in the CPU tests).
"""
import numpy as np


def shard(R, rank, world):
    """Contiguous [lo, hi) slice of R replicas owned by `rank`, sizes differing by at most one."""
    base, rem = divmod(int(R), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_aligned(R, rank, world, align=32):
    """As shard() but every boundary is a multiple of `align` (SA packs 32 restarts per word, SVMC draws one
    Philox call per four reads: their replica_offset must be a multiple of 32 / 4)."""
    blocks = (int(R) + align - 1) // align
    lo, hi = shard(blocks, rank, world)
    return min(lo * align, R), min(hi * align, R)


def gather_best_device(e_local, conf_local, lo, R):
    """Device-resident form of gather_best: `e_local` float64 [n] and `conf_local` int8 [n, N] are torch tensors on
    this rank's GPU (filled by State.best_into).  One all_gather of the padded energy shards, arg-min on the
    device, one broadcast of the winning configuration from its owner (N bytes).  Returns (energies [R] tensor,
    best index int, best_conf tensor) on every rank; no host round trip except the scalar arg-min."""
    import torch
    import torch.distributed as dist

    n = int(e_local.shape[0])
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        b = int(torch.argmin(e_local).item())
        return e_local, b, conf_local[b].clone()
    world, rank = dist.get_world_size(), dist.get_rank()
    bounds = [shard(R, r, world) for r in range(world)]
    nmax = max(h - l for l, h in bounds)
    assert lo == bounds[rank][0] and n == bounds[rank][1] - bounds[rank][0]
    buf = torch.full((nmax,), float("inf"), dtype=torch.float64, device=e_local.device)
    buf[:n] = e_local
    flat_out = torch.empty(world * nmax, dtype=torch.float64, device=e_local.device)
    dist.all_gather_into_tensor(flat_out, buf)  # 1-D concatenation (the form gloo and NCCL both accept)
    out = flat_out.view(world, nmax)
    flat = int(torch.argmin(out).item())  # padding is +inf; ties resolve to the lowest rank, lowest index
    owner, off = divmod(flat, nmax)
    best = bounds[owner][0] + off
    conf = conf_local[off].clone() if rank == owner else torch.empty_like(conf_local[0])
    dist.broadcast(conf, src=owner)
    e = torch.cat([out[r, :bounds[r][1] - bounds[r][0]] for r in range(world)])
    return e, best, conf


def gather_best(local_energy, local_conf, lo, R, device=None):
    """All ranks call this with their shard's best-slice energies (float64 [hi-lo]) and configurations
    ([hi-lo, ...] int8).  Returns (energies [R], best_index, best_conf) on every rank."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        e = np.asarray(local_energy, dtype=np.float64)
        b = int(np.argmin(e))
        return e, b, np.array(local_conf[b])
    world, rank = dist.get_world_size(), dist.get_rank()
    dev = torch.device(device) if device is not None else (
        torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu"))
    bounds = [shard(R, r, world) for r in range(world)]
    nmax = max(h - l for l, h in bounds)
    buf = torch.full((nmax,), float("inf"), dtype=torch.float64, device=dev)
    n = bounds[rank][1] - bounds[rank][0]
    assert lo == bounds[rank][0] and n == len(local_energy)
    buf[:n] = torch.as_tensor(np.asarray(local_energy, dtype=np.float64), device=dev)
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    e = np.concatenate([out[r][:bounds[r][1] - bounds[r][0]].cpu().numpy() for r in range(world)])
    best = int(np.argmin(e))
    owner = next(r for r, (l, h) in enumerate(bounds) if l <= best < h)
    shape = tuple(np.asarray(local_conf).shape[1:])
    conf = torch.empty(shape, dtype=torch.int8, device=dev)
    if rank == owner:
        conf.copy_(torch.as_tensor(np.ascontiguousarray(local_conf[best - bounds[owner][0]]), device=dev))
    dist.broadcast(conf, src=owner)
    return e, best, conf.cpu().numpy()

"""Multi-GPU layout of the path: independent basins (or ensemble members) per rank, one process per
GPU, no halo and no data-path collective.  The only exchange is the gather of the
gauge / outlet time series, done with ``torch.distributed`` (NCCL on the GPUs, gloo in the CPU tests).

Every dependency of the reference follows the LDD downstream and set_dd starts from all pits at once
(wflow/wflow_funcs.py:79), so the trees of the drainage forest never interact; ensemble members are
independent full runs (wflow/wflow_fit.py:235-257).
"""
import numpy as np


def basin_labels(ldd):
    """Label every network cell with the index (1-based, in flat pit order) of the pit it drains to.
    Returns (labels[nrow, ncol] int32 with 0 outside the network, cells_per_basin[int64])."""
    from . import core
    ldd = np.ascontiguousarray(ldd, dtype=np.uint8)
    nodes, up, off = core.build_schedule(ldd)
    labels = np.zeros(ldd.size, np.int32)
    nlev = off.size - 1
    pits = nodes[off[nlev - 1]: off[nlev]] if nlev > 0 else nodes[:0]
    labels[pits] = np.arange(1, pits.size + 1, dtype=np.int32)
    for lev in range(nlev - 1, -1, -1):  # pits first, then one hop upstream at a time
        idx = nodes[off[lev]: off[lev + 1]]
        u = up[off[lev]: off[lev + 1]]
        valid = u >= 0
        labels[u[valid]] = np.repeat(labels[idx], valid.sum(axis=1))
    sizes = np.bincount(labels, minlength=pits.size + 1)[1:].astype(np.int64)
    return labels.reshape(ldd.shape), sizes


def partition(sizes, world_size):
    """Greedy longest-processing-time bin packing of basins (or any independent units) onto ranks by
    cell count.  Returns a list of sorted index arrays (0-based unit indices), one per rank."""
    sizes = np.asarray(sizes, dtype=np.int64)
    order = np.argsort(-sizes, kind="stable")
    load = np.zeros(world_size, np.int64)
    bins = [[] for _ in range(world_size)]
    for i in order:
        r = int(np.argmin(load))
        bins[r].append(int(i))
        load[r] += sizes[i]
    return [np.array(sorted(b), dtype=np.int64) for b in bins]


def shard_ldd(ldd, labels, basins):
    """The LDD restricted to the given basins (0-based indices into the pit order); other cells missing."""
    keep = np.isin(labels, np.asarray(basins) + 1)
    return np.where(keep, ldd, 0).astype(np.uint8)


def shard_members(n_members, rank, world_size):
    """Contiguous block of ensemble members for this rank (members are independent full runs)."""
    per, rem = divmod(n_members, world_size)
    start = rank * per + min(rank, rem)
    return range(start, start + per + (1 if rank < rem else 0))


def gather_series(values, device=None):
    """All-gather a per-rank 1-D float32 vector of (possibly different) length; returns the list of the
    ranks' vectors on every rank.  No-op (single entry) when torch.distributed is not initialised."""
    import torch
    import torch.distributed as dist
    v = torch.as_tensor(np.ascontiguousarray(values, dtype=np.float32))
    if not (dist.is_available() and dist.is_initialized()):
        return [v.numpy().copy()]
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    world = dist.get_world_size()
    n = torch.tensor([v.numel()], dtype=torch.int64, device=device)
    ns = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(ns, n)
    nmax = int(max(int(x.item()) for x in ns))
    buf = torch.zeros(max(nmax, 1), dtype=torch.float32, device=device)
    buf[: v.numel()] = v.to(device)
    out = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    return [o[: int(k.item())].cpu().numpy() for o, k in zip(out, ns)]

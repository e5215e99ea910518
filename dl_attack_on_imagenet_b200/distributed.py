"""Host-side logic of the image-sharded multi-GPU fit (one process per GPU, torch.distributed / NCCL).

ADiL shards naturally: images -- and with them the code rows v[N_r, K] and their AdamW state -- are partitioned
contiguously by rank; the dictionary D is replicated.  The only data-path collective is one SUM all-reduce of
dD [P, K] per minibatch step (the reference loss is CrossEntropy(reduction='sum'), adil.py:136, so summing the
per-rank gradients reproduces the single-GPU gradient of the union batch).  These helpers are pure functions so
that the schedule / partition logic is testable on CPU with the gloo backend.
"""
import torch


def shard_bounds(n, world, rank):
    """Contiguous near-equal partition of range(n): the first n % world ranks get one extra item."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def owner_of(index, n, world):
    """Rank owning global row `index` under shard_bounds."""
    base, extra = divmod(n, world)
    cut = extra * (base + 1)
    if index < cut:
        return index // (base + 1)
    return extra + (index - cut) // max(base, 1)


def schedule_seed(obj):
    return int(getattr(obj, 'schedule_seed', 0))


def epoch_schedule(n, world, batch_size, epoch, seed=0):
    """Deterministic minibatch schedule of one epoch, identical on every rank.

    Returns a list over steps; each step is a list over ranks of int64 CPU tensors holding the GLOBAL indices
    (all inside that rank's shard) processed at that step.  Every rank shuffles its own shard with a generator
    seeded by (seed, epoch, rank); ranks whose shard is exhausted get an empty tensor."""
    perms = []
    for r in range(world):
        lo, hi = shard_bounds(n, world, r)
        g = torch.Generator().manual_seed((seed * 1000003 + epoch) * 8191 + r)
        perms.append(lo + torch.randperm(hi - lo, generator=g))
    n_steps = max((len(p) + batch_size - 1) // batch_size for p in perms) if n > 0 else 0
    return [[p[t * batch_size:(t + 1) * batch_size] for p in perms] for t in range(n_steps)]


def union_schedule(n, world, batch_size, epoch, seed=0):
    """The same schedule seen by ONE process: per step the concatenation of all ranks' indices (used to check that
    R ranks x B images == one GPU with batch R*B)."""
    return [torch.cat(step) for step in epoch_schedule(n, world, batch_size, epoch, seed)]


def allreduce_sum_(t):
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def gather_rows(local_rows, n, world, rank):
    """All-gather row shards of unequal size into the full [n, K] tensor (every rank gets it)."""
    import torch.distributed as dist
    K = local_rows.shape[1]
    base, extra = divmod(n, world)
    cap = base + (1 if extra else 0)
    pad = torch.zeros(cap, K, dtype=local_rows.dtype, device=local_rows.device)
    pad[:local_rows.shape[0]] = local_rows
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad)
    parts = []
    for r in range(world):
        lo, hi = shard_bounds(n, world, r)
        parts.append(bufs[r][:hi - lo])
    return torch.cat(parts, dim=0)

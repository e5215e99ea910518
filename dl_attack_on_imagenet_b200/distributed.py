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


# ----------------------------------------------------------------------------------------------------------
# sharded dictionary step: reduce-scatter -> AdamW + clamp on this rank's pixel slice -> all-gather
# ----------------------------------------------------------------------------------------------------------
def padded_rows(P, world):
    """Rows per rank (and padded total) of the [P, K] dictionary split into `world` equal pixel slices."""
    per = (P + world - 1) // world
    return per, per * world


class ShardedDictStep(object):
    """The dictionary side of one multi-GPU minibatch step (intent of adil.py:379-383):

        dD2 (this rank's gradient, [P, K])  --reduce-scatter (SUM)-->  dD of pixel slice r
        AdamW + clamp on slice r of D2 (moments m, s exist ONLY for that slice: optimizer state sharded R-fold)
        slice r of D2                        --all-gather-->            D2 on every rank

    Same wire bytes as the all-reduce it replaces (2 (R-1)/R x 4PK), but the AdamW pass and its HBM traffic drop from
    28 PK bytes per rank to 28 PK / R.  SUM, not mean: the reference loss is CrossEntropy(reduction='sum').

    `D2` and `dD2` must be the first P rows of buffers with `rows_total` rows (see `alloc`), so that every rank's slice
    has the same size.  `step_fn(D_slice, m, s, dD_slice, hp, atoms_mode)` defaults to the CUDA kernel
    (ops.dict_step); the gloo/CPU test injects a host implementation there.  The collectives run on `stream` (a side stream: the
    local code step proceeds concurrently); `wait()` makes the current stream wait for the gathered dictionary."""

    def __init__(self, P, K, device, group=None, step_fn=None, side_stream=True):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.P, self.K = P, K
        self.rows, self.rows_total = padded_rows(P, self.world)
        self.device = torch.device(device)
        self.lo = self.rank * self.rows
        self.m = torch.zeros(self.rows, K, device=self.device)
        self.s = torch.zeros(self.rows, K, device=self.device)
        self.dD_slice = torch.empty(self.rows, K, device=self.device)
        self.t = 0
        self.backend = dist.get_backend(group)
        self.stream = None
        self._done = None
        if self.device.type == 'cuda' and side_stream:
            self.stream = torch.cuda.Stream(device=self.device)
        if step_fn is None:
            from . import ops

            def step_fn(D_slice, m, s, dD_slice, hp, atoms_mode):
                ops.dict_step(D_slice, m, s, dD_slice, hp, atoms_mode)
        self.step_fn = step_fn

    def alloc(self):
        """Zero-filled [rows_total, K] buffer; its first P rows are the usable [P, K] view."""
        return torch.zeros(self.rows_total, self.K, device=self.device)

    def _reduce_scatter(self, full):
        dist = self.dist
        if self.world == 1:
            self.dD_slice.copy_(full)
        elif self.backend == 'nccl':
            dist.reduce_scatter_tensor(self.dD_slice, full, op=dist.ReduceOp.SUM, group=self.group)
        else:  # gloo has no reduce-scatter: all-reduce and keep the slice
            dist.all_reduce(full, op=dist.ReduceOp.SUM, group=self.group)
            self.dD_slice.copy_(full[self.lo:self.lo + self.rows])

    def _all_gather(self, full):
        dist = self.dist
        mine = full[self.lo:self.lo + self.rows]
        if self.world == 1:
            return
        if self.backend == 'nccl':
            dist.all_gather_into_tensor(full, mine, group=self.group)   # in place: slice r sits at offset r*rows
        else:
            parts = [torch.empty_like(mine) for _ in range(self.world)]
            dist.all_gather(parts, mine.clone(), group=self.group)
            for r, p in enumerate(parts):
                full[r * self.rows:(r + 1) * self.rows].copy_(p)

    def step(self, D_full, dD_full, hp, atoms_mode):
        """D_full, dD_full: the [rows_total, K] buffers.  Enqueues reduce-scatter, slice step and all-gather."""
        if self.stream is not None:
            self.stream.wait_stream(torch.cuda.current_stream(self.device))
            ctx = torch.cuda.stream(self.stream)
        else:
            import contextlib
            ctx = contextlib.nullcontext()
        with ctx:
            self._reduce_scatter(dD_full)
            self.t += 1
            self.step_fn(D_full[self.lo:self.lo + self.rows], self.m, self.s, self.dD_slice, hp, atoms_mode)
            self._all_gather(D_full)
            if self.stream is not None:
                self._done = torch.cuda.Event()
                self._done.record(self.stream)

    def wait(self):
        """The current stream waits until the dictionary of the last `step` is complete on this rank."""
        if self._done is not None:
            torch.cuda.current_stream(self.device).wait_event(self._done)
            self._done = None


class PeerDictStep(ShardedDictStep):
    """The same sharded dictionary step as ONE kernel over peer-mapped memory (adil_dict_step_peer): the dictionary and
    the gradient buffers are allocated as torch symmetric memory (every rank's buffer mapped into every process over
    NVLink / NVSwitch), and per step

        cross-rank barrier           every rank's adil_grad has written its dD buffer
        adil_dict_step_peer          peer loads of the R gradient slices, AdamW + clamp, peer stores of the new slice
        cross-rank barrier           every rank's stores have landed before D is read again

    replace reduce-scatter -> slice step -> all-gather (three launches through NCCL, two staging passes).  PyTorch only
    provides the mapping and the barrier; the data movement is the kernel's own loads and stores -- through the
    NVSwitch's multicast address when the allocation has one (multimem.ld_reduce / multimem.st: the switch adds the R
    gradient values in flight and replicates the result), else R peer loads and R peer stores per element."""

    def __init__(self, P, K, device, group=None, side_stream=True):
        super().__init__(P, K, device, group=group, step_fn=None, side_stream=side_stream)
        import torch.distributed._symmetric_memory as symm_mem
        self.symm_mem = symm_mem
        self.pg = group if group is not None else self.dist.group.WORLD
        self._handles = {}
        self._bufs = []
        import os
        # ADIL_DICT_STEP_MULTICAST=1: NVLS (multicast + in-switch reduction) when the symmetric allocation carries a
        # multicast address.  Off by default: measured slower than the plain peer loads / stores at 2 ranks (115 vs 66
        # us at K = 50), where the switch saves no traffic; scripts/dist_parity.py times both.
        self.use_multicast = os.environ.get("ADIL_DICT_STEP_MULTICAST", "0") == "1"

    def alloc(self):
        t = self.symm_mem.empty(self.rows_total * self.K, dtype=torch.float32, device=self.device)
        hdl = self.symm_mem.rendezvous(t, self.pg)
        t.zero_()
        self._handles[t.data_ptr()] = hdl
        self._bufs.append(t)
        return t.view(self.rows_total, self.K)

    def step(self, D_full, dD_full, hp, atoms_mode):
        from . import ops
        hD, hG = self._handles[D_full.data_ptr()], self._handles[dD_full.data_ptr()]
        if self.stream is not None:
            self.stream.wait_stream(torch.cuda.current_stream(self.device))
            ctx = torch.cuda.stream(self.stream)
        else:
            import contextlib
            ctx = contextlib.nullcontext()
        with ctx:
            hG.barrier(channel=0, timeout_ms=20000)   # (bounded: a protocol bug traps instead of hanging the GPU)
            self.t += 1
            mc_D, mc_G = (int(hD.multicast_ptr or 0), int(hG.multicast_ptr or 0)) if self.use_multicast else (0, 0)
            ops.dict_step_peer(list(hD.buffer_ptrs), list(hG.buffer_ptrs), self.m, self.s, self.lo * self.K,
                               self.rows * self.K, self.rank, hp, atoms_mode, device=self.device,
                               D_mc=mc_D if (mc_D and mc_G) else 0, dD_mc=mc_G if (mc_D and mc_G) else 0)
            hD.barrier(channel=1, timeout_ms=20000)
            if self.stream is not None:
                self._done = torch.cuda.Event()
                self._done.record(self.stream)


def make_dict_step(P, K, device, group=None, mode=None):
    """The sharded dictionary step for this process group: the fused peer-memory kernel when symmetric memory can be
    set up on every rank (NCCL group, CUDA device, world > 1; ADIL_DICT_STEP=nccl forces the NCCL path), else
    reduce-scatter / slice step / all-gather through NCCL.  Collective: every rank must call it."""
    import os
    import torch.distributed as dist
    mode = mode or os.environ.get("ADIL_DICT_STEP", "auto")
    world = dist.get_world_size(group)
    dev = torch.device(device)
    if mode != "nccl" and world > 1 and dev.type == 'cuda' and dist.get_backend(group) == 'nccl':
        step, ok = None, 1
        try:
            step = PeerDictStep(P, K, dev, group=group)
            probe = step.alloc()                              # symmetric allocation + rendezvous work on this rank
            step._handles.pop(probe.data_ptr(), None)
            step._bufs.clear()
            del probe
        except Exception:
            ok = 0
        flag = torch.tensor([ok], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)   # all ranks take the same path
        if int(flag.item()) == 1:
            return step
        if mode == "peer":
            raise RuntimeError("ADIL_DICT_STEP=peer: symmetric memory could not be set up on every rank")
    return ShardedDictStep(P, K, dev, group=group)

"""Host-side logic that needs no GPU: sharding / schedules, Normalize peeling, AdamW scalar packing, and the
world_size-2 gloo run of the image-sharded step (kernels replaced by the oracle -- test infrastructure only)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dl_attack_on_imagenet_b200 import distributed as dsh
from dl_attack_on_imagenet_b200 import ops, split_normalize
from dl_attack_on_imagenet_b200.data import IndexedTensorDataset, Normalize
from oracle import adil_oracle as O


def test_shard_bounds_partition():
    for n in (0, 1, 7, 32, 1000, 16384):
        for world in (1, 2, 3, 8):
            spans = [dsh.shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for a, b in zip(spans[:-1], spans[1:]):
                assert a[1] == b[0]
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
            for i in range(0, n, max(1, n // 17)):
                r = dsh.owner_of(i, n, world)
                assert spans[r][0] <= i < spans[r][1]


def test_epoch_schedule_covers_each_shard_once():
    n, world, bs = 37, 4, 5
    sched = dsh.epoch_schedule(n, world, bs, epoch=3, seed=11)
    seen = torch.cat([torch.cat(step) for step in sched])
    assert sorted(seen.tolist()) == list(range(n))
    for step in sched:
        for r, idx in enumerate(step):
            lo, hi = dsh.shard_bounds(n, world, r)
            assert idx.numel() <= bs and all(lo <= i < hi for i in idx.tolist())
    again = dsh.epoch_schedule(n, world, bs, epoch=3, seed=11)
    assert all(torch.equal(a, b) for sa, sb in zip(sched, again) for a, b in zip(sa, sb))
    other = dsh.epoch_schedule(n, world, bs, epoch=4, seed=11)
    assert any(not torch.equal(a, b) for sa, sb in zip(sched, other) for a, b in zip(sa, sb))
    uni = dsh.union_schedule(n, world, bs, epoch=3, seed=11)
    assert all(torch.equal(u, torch.cat(s)) for u, s in zip(uni, sched))


def test_split_normalize():
    net = torch.nn.Sequential(torch.nn.Conv2d(3, 4, 3), torch.nn.Flatten())
    model = torch.nn.Sequential(Normalize(), net)
    rest, mean, std = split_normalize(model)
    assert rest is net
    assert mean == pytest.approx([0.485, 0.456, 0.406]) and std == pytest.approx([0.229, 0.224, 0.225])
    rest2, mean2, _ = split_normalize(net)
    assert rest2 is net and mean2 is None
    x = torch.rand(2, 3, 8, 8)
    assert torch.equal(model(x), rest((x - torch.tensor(mean).view(1, 3, 1, 1)) / torch.tensor(std).view(1, 3, 1, 1)))


def test_adamw_params_struct():
    hp = ops.adamw_params(3, 0.01)
    assert (hp.lr, hp.beta1, hp.beta2, hp.eps, hp.weight_decay, hp.step) == (0.01, 0.9, 0.999, 1e-8, 1e-2, 3)


def test_indexed_dataset_protocol():
    ds = IndexedTensorDataset(torch.rand(5, 3, 4, 4), torch.arange(5))
    x, y = ds[2]
    assert x.shape == (3, 4, 4) and int(y) == 2
    ds.indexed = True
    i, x, y = ds[3]
    assert i == 3 and int(y) == 3


# ---- world_size = 2 over gloo: R ranks x B images == one process with the union batch ----------------------
C, H, W, K, N, B = 3, 8, 8, 5, 12, 3
EPS = 8.0 / 255.0


def _problem():
    g = torch.Generator().manual_seed(21)
    D = -1 + 2 * torch.rand(C, H, W, K, generator=g)
    v = O.project_rows_l1(torch.rand(N, K, generator=g), EPS)
    x = torch.rand(N, C, H, W, generator=g)
    return D, v, x


def _oracle_slice_step(D_slice, m, s, dD_slice, hp, atoms_mode):
    """AdamW + clamp on a pixel slice (what ops.dict_step does on the GPU) -- test infrastructure."""
    O.adamw_step_(D_slice, dD_slice, m, s, hp.step, hp.lr)
    if atoms_mode == ops.ATOMS_CLAMP1:
        D_slice.clamp_(-1, 1)


def _sharded_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    model = O.tiny_classifier(seed=0)
    net, mean, std = O.split_normalize(model)
    D, v, x = _problem()
    P = C * H * W
    lo, hi = dsh.shard_bounds(N, world, rank)
    st = O.State(D, v[lo:hi])                      # each rank owns its images' code rows only
    # the product's sharded dictionary step (reduce-scatter -> slice AdamW -> all-gather), oracle arithmetic on CPU
    shard = dsh.ShardedDictStep(P, K, 'cpu', step_fn=_oracle_slice_step)
    D_full, dD_full = shard.alloc(), shard.alloc()
    D_full[:P].copy_(st.D2)
    st.D2 = D_full[:P]
    for epoch in range(2):
        for step in dsh.epoch_schedule(N, world, B, epoch, seed=5):
            idx = step[rank]
            if idx.numel():
                with torch.no_grad():
                    labels = model(x[idx]).argmax(-1)
                xin, _ = O.synth(x[idx].reshape(len(idx), P), st.D2, st.v, idx - lo, mean, std, EPS, O.F_NORMALIZE)
                _, g, _ = O.classifier_grad(net, xin.reshape(-1, C, H, W), labels, 'ce', 50, False, 'sum')
                dD2, dvb = O.grad(g.reshape(len(idx), P), st.D2, st.v[idx - lo], std)
            else:
                dD2, dvb = torch.zeros_like(st.D2), torch.zeros(0, K)
            dD_full.zero_()
            dD_full[:P].copy_(dD2)
            st.tD += 1
            shard.step(D_full, dD_full, ops.adamw_params(st.tD, 0.01), ops.ATOMS_CLAMP1)   # the data-path collectives
            shard.wait()
            O.code_step_(st, dvb, idx - lo, 0.01, EPS)
    v_all = dsh.gather_rows(st.v, N, world, rank)
    if rank == 0:
        torch.save({"D2": st.D2.clone(), "v": v_all, "m_rows": shard.m.shape[0]}, os.path.join(out_dir, "sharded.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_gloo_ranks_match_single_process(tmp_path, world):
    port = 29500 + (os.getpid() % 2000) + 37 * world
    mp.spawn(_sharded_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    got = torch.load(os.path.join(str(tmp_path), "sharded.pt"))
    # single process, union batches
    torch.set_num_threads(1)
    model = O.tiny_classifier(seed=0)
    net, mean, std = O.split_normalize(model)
    D, v, x = _problem()
    P = C * H * W
    st = O.State(D, v)
    for epoch in range(2):
        for idx in dsh.union_schedule(N, world, B, epoch, seed=5):
            with torch.no_grad():
                labels = model(x[idx]).argmax(-1)
            xin, _ = O.synth(x[idx].reshape(len(idx), P), st.D2, st.v, idx, mean, std, EPS, O.F_NORMALIZE)
            _, g, _ = O.classifier_grad(net, xin.reshape(-1, C, H, W), labels, 'ce', 50, False, 'sum')
            O.joint_step_(st, g.reshape(len(idx), P), idx, 0.01, EPS, std)
    assert (got["D2"] - st.D2).abs().max() < 2e-6
    assert (got["v"] - st.v).abs().max() < 2e-6
    assert got["m_rows"] == (C * H * W + world - 1) // world       # optimizer state sharded R-fold


def _padded_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    P, K = 100, 7                                   # 100 rows over 3 ranks: slices of 34 rows, 2 rows of padding
    shard = dsh.ShardedDictStep(P, K, 'cpu', step_fn=_oracle_slice_step)
    g = torch.Generator().manual_seed(3)
    D0 = -1 + 2 * torch.rand(P, K, generator=g)
    grads = [torch.randn(world, P, K, generator=g) * 1e-3 for _ in range(3)]
    D_full, dD_full = shard.alloc(), shard.alloc()
    D_full[:P].copy_(D0)
    for t, gr in enumerate(grads):
        dD_full[:P].copy_(gr[rank])
        shard.step(D_full, dD_full, ops.adamw_params(t + 1, 0.01), ops.ATOMS_CLAMP1)
        shard.wait()
    if rank == 1:
        torch.save({"D": D_full.clone(), "rows": (shard.rows, shard.rows_total)}, os.path.join(out_dir, "padded.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_dict_step_with_padded_slices(tmp_path):
    world = 3
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_padded_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    got = torch.load(os.path.join(str(tmp_path), "padded.pt"))
    assert got["rows"] == (34, 102)
    P, K = 100, 7
    g = torch.Generator().manual_seed(3)
    D = -1 + 2 * torch.rand(P, K, generator=g)
    grads = [torch.randn(world, P, K, generator=g) * 1e-3 for _ in range(3)]
    m, s = torch.zeros_like(D), torch.zeros_like(D)
    for t, gr in enumerate(grads):
        O.adamw_step_(D, gr.sum(0), m, s, t + 1, 0.01)
        D.clamp_(-1, 1)
    assert (got["D"][:P] - D).abs().max() < 2e-6 and (got["D"][P:] == 0).all()


# ---- transfer sweep (performance.py:183-232), sharded over 2 gloo ranks: the counters are all-reduced ------------------
class _ShiftAttack(object):
    device = torch.device("cpu")

    def __call__(self, x, y):
        return (x + 0.2 * torch.sign(x - 0.5)).clamp(0, 1)


def _host_errors(adv, clean):
    d = (adv - clean).flatten(1)
    return (d ** 2).sum(1), (clean.flatten(1) ** 2).sum(1), d.abs().amax(1)


def _sweep_models():
    torch.manual_seed(5)
    return {"a": torch.nn.Sequential(torch.nn.Flatten(), torch.nn.Linear(3 * 8 * 8, 7)),
            "b": torch.nn.Sequential(torch.nn.Flatten(), torch.nn.Linear(3 * 8 * 8, 7))}


def _sweep_data():
    g = torch.Generator().manual_seed(6)
    x = torch.rand(22, 3, 8, 8, generator=g)
    y = torch.randint(0, 7, (22,), generator=g)
    return x, y


def _sweep_worker(rank, world, port, out_dir):
    from dl_attack_on_imagenet_b200 import performance as perf
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    x, y = _sweep_data()
    lo, hi = dsh.shard_bounds(len(x), world, rank)                       # unequal shards: 11 + 11, batches of 4
    data = [(x[i:min(i + 4, hi)], y[i:min(i + 4, hi)]) for i in range(lo, hi, 4)]
    res = perf.get_transfer_performance({"adil": [_ShiftAttack()], "none": []}, _sweep_models(), data,
                                        errors_fn=_host_errors)
    if rank == 1:
        torch.save(res, os.path.join(out_dir, "sweep.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_transfer_sweep_counters_are_summed_over_ranks(tmp_path):
    from dl_attack_on_imagenet_b200 import performance as perf
    world = 2
    port = 33500 + (os.getpid() % 2000)
    mp.spawn(_sweep_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    got = torch.load(os.path.join(str(tmp_path), "sweep.pt"))
    x, y = _sweep_data()
    ref = perf.get_transfer_performance({"adil": [_ShiftAttack()], "none": []}, _sweep_models(), [(x, y)],
                                        errors_fn=_host_errors)            # one process, the whole set
    for name in ("a", "b"):
        for key in ("fooling_rate", "rmse", "mse"):
            assert got["adil"][name][key] == pytest.approx(ref["adil"][name][key], rel=1e-6)
        assert got["none"][name]["mse"] != got["none"][name]["mse"]       # NaN, performance.py:198-202
    # against the reference's formulas (performance.py:238-266)
    adv = _ShiftAttack()(x, y)
    m = _sweep_models()["b"]
    assert got["adil"]["b"]["fooling_rate"] == pytest.approx((m(x).argmax(1) != m(adv).argmax(1)).float().mean().item(), rel=1e-6)
    assert got["adil"]["a"]["mse"] == pytest.approx(((adv - x) ** 2).sum(dim=[1, 2, 3]).mean().item(), rel=1e-6)

"""Multi-GPU parity of the PRODUCT distributed fit, run under `gpurun --gpus 2 -- torchrun ... scripts/dist_parity.py`.

ADIL.learn_dictionary_distributed on R ranks (image-sharded; dD reduce-scatter -> AdamW on each rank's pixel slice ->
all-gather of D; code rows local) against the single-GPU fused step on the UNION batch, teacher-forced step by step:
before every minibatch step the full state of the distributed run (D, the gathered moment slices, all code rows and
their moments) is snapshotted; after it rank 0 replays the same step on one GPU -- fused adil_grad_dict_step +
adil_code_step on the concatenated batch with the classifier gradients the ranks actually used -- and compares the
post-step states.  Bounds: m within 1e-6 relative, v within 2e-6; D is compared where AdamW is
well conditioned and as a fraction of entries.  Also times the pieces of the sharded step and
the NCCL collectives on the dictionary-gradient payloads of the BASELINE configs.

Writes gpurun_out/dist_parity_r02.json (rank 0).
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    from dl_attack_on_imagenet_b200 import ADIL, IndexedTensorDataset, build_classifier, ops, synthetic_images
    from dl_attack_on_imagenet_b200 import distributed as dsh
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    EPS = 8.0 / 255.0
    N, K, B, EPOCHS = 32 * world, int(os.environ.get("ADIL_PARITY_K", "50")), 16, 2   # (K = 200: the column-window kernels)
    arch = os.environ.get("ADIL_PARITY_MODEL", "resnet18")
    model = build_classifier(arch, seed=0, device=dev)
    x, y = synthetic_images(N, seed=1)
    data = IndexedTensorDataset(x, y)
    ADIL.verbose = False
    P = 3 * 224 * 224
    std = [0.229, 0.224, 0.225]
    report = {"world": world, "model": arch, "N": N, "K": K, "batch_per_rank": B, "steps": []}
    stash = {}
    orig_cg = ADIL._classifier_grad
    orig_step = ADIL._fit_step

    def classifier_grad(self, xin, labels, reduction):
        loss, g, out = orig_cg(self, xin, labels, reduction)
        stash["g"] = g.detach().reshape(g.shape[0], -1).clone()
        return loss, g, out

    def gather_cat(t):
        """Concatenate equally-shaped per-rank tensors along dim 0 (on every rank)."""
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t.contiguous())
        return torch.cat(parts, dim=0)

    def snapshot(st):
        st.shard.wait()
        torch.cuda.synchronize()
        rows = st.shard.rows
        return {"D2": st.D2.clone(), "m": gather_cat(st.shard.m)[:P], "s": gather_cat(st.shard.s)[:P],
                "v": gather_cat(st.v), "mv": gather_cat(st.mv), "sv": gather_cat(st.sv), "rows": rows}

    def fit_step(self, st, x_src, x_index, v_index, labels, shape, update, lr_d, lr_v, index_cpu=None):
        lo, _ = dsh.shard_bounds(N, world, rank)
        pre = snapshot(st)
        tD, tv = st.tD + 1, st.tv + 1
        loss, fooled = orig_step(self, st, x_src, x_index, v_index, labels, shape, update, lr_d, lr_v, index_cpu=index_cpu)
        post = snapshot(st)
        g_all = gather_cat(stash["g"])                                   # [world*B, P] in rank order
        idx_all = gather_cat((index_cpu + lo).to(dev))                   # global rows of v, rank order
        # every rank holds the same replicated dictionary after the all-gather
        d_rep = post["D2"].clone()
        dist.broadcast(d_rep, 0)
        rep_gap = (d_rep - post["D2"]).abs().max().item()
        if rank == 0:
            D, m, s = pre["D2"].clone(), pre["m"].clone(), pre["s"].clone()
            v, mv, sv = pre["v"].clone(), pre["mv"].clone(), pre["sv"].clone()
            part = ops.grad_dict_step(D, m, s, g_all, v, idx_all.cpu(), ops.adamw_params(tD, lr_d), std,
                                      ops.ATOMS_CLAMP1, keep_partials=True)
            ops.code_step(v, mv, sv, part, idx_all, ops.adamw_params(tv, lr_v), ops.ROWS_L1BALL, EPS)
            torch.cuda.synchronize()
            dD_single = (m.double() - 0.9 * pre["m"].double()) / 0.1       # the gradient each path applied
            dD_dist = (post["m"].double() - 0.9 * pre["m"].double()) / 0.1
            big = dD_single.abs() > 1e-3 * dD_single.abs().max()
            dgap = (D - post["D2"]).abs()
            rec = {"t": tD,
                   "m_rel": ((m - post["m"]).abs().max() / post["m"].abs().max()).item(),
                   "s_rel": ((s - post["s"]).abs().max() / post["s"].abs().max()).item(),
                   "dD_rel": ((dD_single - dD_dist).abs().max() / dD_single.abs().max()).item(),
                   "v_abs": (v - post["v"]).abs().max().item(),
                   "mv_rel": ((mv - post["mv"]).abs().max() / post["mv"].abs().max().clamp_min(1e-30)).item(),
                   "D_abs_where_conditioned": (dgap * big).max().item(),
                   "D_abs_max": dgap.max().item(), "D_frac_gt_1e-6": (dgap > 1e-6).float().mean().item(),
                   "replicas_gap": rep_gap, "loss": float(loss)}
            report["steps"].append(rec)
        dist.barrier()
        return loss, fooled

    ADIL._classifier_grad = classifier_grad
    ADIL._fit_step = fit_step
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    os.chdir(os.path.join(ROOT, "gpurun_out"))
    stale = os.path.join("trained_dicts", "ImageNet_dist_parity_%d.bin" % rank)
    if os.path.exists(stale):                    # (a saved dictionary would skip the fit, adil.py:94)
        os.remove(stale)
    dist.barrier()
    torch.manual_seed(1234)
    atk = ADIL(model, eps=EPS, steps=EPOCHS, n_atoms=K, batch_size=B, data_train=data, model_name="dist_parity_%d" % rank,
               is_distributed=True, loss='ce', method='gd')
    ADIL._classifier_grad = orig_cg
    ADIL._fit_step = orig_step
    st = atk.state
    ok = True
    if rank == 0:
        worst = {k: max(r[k] for r in report["steps"]) for k in report["steps"][0] if k not in ("t", "loss")}
        report["worst"] = worst
        ok = (worst["m_rel"] <= 1e-6 and worst["v_abs"] <= 2e-6 and worst["D_abs_where_conditioned"] <= 1e-6 and
              worst["D_frac_gt_1e-6"] <= 1e-3 and worst["replicas_gap"] == 0.0)
        report["pass"] = bool(ok)
        report["optimizer_state_rows_per_rank"] = int(st.shard.rows)
        report["dict_step_backend"] = type(st.shard).__name__
        if isinstance(st.shard, dsh.PeerDictStep):
            h0 = next(iter(st.shard._handles.values()))
            report["dict_step_multicast"] = bool(st.shard.use_multicast and int(h0.multicast_ptr or 0))

    # ---- timing: pieces of the sharded step and the collectives at the BASELINE payloads -----------------------------
    def timed(fn, iters=20, warm=5):
        for _ in range(warm):
            fn()
        dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters):
            fn()
        b.record()
        torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b) / iters], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item() * 1e3                                             # us, max over ranks

    times = {}
    for Kc in (50, 64, 100, 200):
        nbytes = 4 * P * Kc
        rows, total = dsh.padded_rows(P, world)
        full = torch.randn(total, Kc, device=dev)
        sl = torch.empty(rows, Kc, device=dev)
        Df, ms_, ss_ = torch.rand(total, Kc, device=dev), torch.zeros(rows, Kc, device=dev), torch.zeros(rows, Kc, device=dev)
        hp = ops.adamw_params(3, 0.01)
        lo_r = rank * rows

        def rs():
            dist.reduce_scatter_tensor(sl, full, op=dist.ReduceOp.SUM)

        def ag():
            dist.all_gather_into_tensor(Df, Df[lo_r:lo_r + rows])

        def ar():
            dist.all_reduce(full, op=dist.ReduceOp.SUM)

        def sharded():
            rs()
            ops.dict_step(Df[lo_r:lo_r + rows], ms_, ss_, sl, hp, ops.ATOMS_CLAMP1)
            ag()

        mf, sf = torch.zeros(total, Kc, device=dev), torch.zeros(total, Kc, device=dev)

        def replicated():
            ar()
            ops.dict_step(Df, mf, sf, full, hp, ops.ATOMS_CLAMP1)

        entry = {
            "reduce_scatter_us": timed(rs), "all_gather_us": timed(ag), "all_reduce_us": timed(ar),
            "sharded_step_us (rs + slice AdamW + ag)": timed(sharded),
            "replicated_step_us (all-reduce + full AdamW)": timed(replicated)}
        if isinstance(st.shard, dsh.PeerDictStep):
            peer = dsh.PeerDictStep(P, Kc, dev, side_stream=False)
            Dp, Gp = peer.alloc(), peer.alloc()
            Dp.uniform_(-1, 1)
            Gp.normal_()

            def fused():
                peer.step(Dp, Gp, hp, ops.ATOMS_CLAMP1)
            hD, hG = peer._handles[Dp.data_ptr()], peer._handles[Gp.data_ptr()]
            mc = (int(hD.multicast_ptr or 0), int(hG.multicast_ptr or 0))
            entry["multicast_available"] = bool(mc[0] and mc[1])
            peer.use_multicast = False
            entry["peer_step_us (barrier + one fused kernel over NVLink + barrier)"] = timed(fused)

            def kernel_only():
                ops.dict_step_peer(list(hD.buffer_ptrs), list(hG.buffer_ptrs), peer.m, peer.s, peer.lo * Kc, peer.rows * Kc,
                                   rank, hp, ops.ATOMS_CLAMP1, device=dev)
            entry["peer_kernel_only_us"] = timed(kernel_only)
            if mc[0] and mc[1]:
                peer.use_multicast = True
                entry["multimem_step_us (barrier + one NVLS kernel: multimem.ld_reduce / multimem.st + barrier)"] = timed(fused)

                def mc_kernel_only():
                    ops.dict_step_peer(list(hD.buffer_ptrs), list(hG.buffer_ptrs), peer.m, peer.s, peer.lo * Kc,
                                       peer.rows * Kc, rank, hp, ops.ATOMS_CLAMP1, device=dev, D_mc=mc[0], dD_mc=mc[1])
                entry["multimem_kernel_only_us"] = timed(mc_kernel_only)
            # the fused kernel against the NCCL chain on the same inputs: same sum, same AdamW
            Dn = Dp.clone()
            mn, sn = torch.zeros(rows, Kc, device=dev), torch.zeros(rows, Kc, device=dev)
            peer.m.zero_(); peer.s.zero_()
            Gn = Gp.clone()
            dist.barrier()
            peer.step(Dp, Gp, hp, ops.ATOMS_CLAMP1)
            dist.reduce_scatter_tensor(sl, Gn, op=dist.ReduceOp.SUM)
            ops.dict_step(Dn[lo_r:lo_r + rows], mn, sn, sl, hp, ops.ATOMS_CLAMP1)
            dist.all_gather_into_tensor(Dn, Dn[lo_r:lo_r + rows])
            torch.cuda.synchronize()
            entry["peer_vs_nccl_m_rel"] = ((peer.m - mn).abs().max() / mn.abs().max()).item()
            entry["peer_vs_nccl_D_frac_gt_1e-6"] = ((Dp - Dn).abs() > 1e-6).float().mean().item()
            del peer, Dp, Gp, Dn, Gn
        times["K=%d (%.1f MB)" % (Kc, nbytes / 1e6)] = entry
        del full, sl, Df, ms_, ss_, mf, sf
    if rank == 0:
        report["collective_times_max_over_ranks"] = times
        tag = os.environ.get("ADIL_DICT_STEP", "auto")
        with open(os.path.join(ROOT, "gpurun_out", "dist_parity_r02_w%d_%s%s.json" % (world, tag, "" if K == 50 else "_k%d" % K)), "w") as f:
            json.dump(report, f, indent=1)
        print(json.dumps({"backend": report["dict_step_backend"], "multicast": report.get("dict_step_multicast"),
                          "pass": report["pass"], "worst": report["worst"],
                          "times": times}, indent=1))
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0 and not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py -- ADiL attack-learning throughput on B200 (BASELINE.json metric: attack images/sec; fused-kernel HBM
GB/s vs peak).

    python bench.py --gpus 1 --steps K --warmup W            # our arm (CUDA kernels behind the C ABI)
    python bench.py --impl reference --steps K --warmup W    # reference arm: the oracle port of the reference's
                                                             # own PyTorch-CPU path on the box's host cores
    torchrun --nproc-per-node N bench.py --gpus N ...        # image-sharded weak scaling, dD all-reduce over NCCL

A "step" is one minibatch of the joint dictionary/code update (adil.py:168-188): clean forward for the labels,
perturbation synthesis, classifier forward + backward, fused backward contractions + AdamW(D) + clamp, code AdamW
+ l1 projection.  Workload at N=1: BASELINE.json configs[1] -- random-init ResNet-50, 1024 synthetic 3x224x224
images, 50 atoms, batch 100, l_inf eps=8/255, fp32.  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

P_IMG = 3 * 224 * 224
EPS = 8.0 / 255.0


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="resnet50")
    ap.add_argument("--atoms", type=int, default=50)
    ap.add_argument("--images", type=int, default=1024, help="images per GPU (weak scaling)")
    ap.add_argument("--batch", type=int, default=100, help="minibatch per GPU")
    ap.add_argument("--ref-batch", type=int, default=32, help="images per step of the CPU reference sample")
    ap.add_argument("--tf32", action="store_true", help="allow TF32 in the cuDNN classifier (default: strict fp32)")
    ap.add_argument("--kernel-impl", default="auto", choices=["auto", "fma", "tc"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cudnn-autotune", action="store_true",
                    help="cudnn.benchmark off (profiling runs: fewer trial kernels in the launch list)")
    return ap.parse_args()


def measured_traffic(kernel):
    """DRAM bytes per launch of `kernel` from the committed ncu --set full capture (profiles/ncu_traffic.json):
    dram__bytes_read.sum + dram__bytes_write.sum.  None when no capture is recorded for it."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        with open(path) as f:
            d = json.load(f)[kernel]
        return float(d["dram_bytes_read"]) + float(d["dram_bytes_write"])
    except Exception:
        return None


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            with open(path) as f:
                d = json.load(f)
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU during the timed region (pynvml, 100 ms period)."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples, self.reasons = [], set()
        self.sm_max = None
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        while not self._halt.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                try:
                    mask = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._halt.wait(0.1)

    def finish(self):
        self._halt.set()
        if self.is_alive():
            self.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons)}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons)}


def make_images(n, seed, pin):
    g = torch.Generator().manual_seed(seed)
    x = torch.empty(n, 3, 224, 224, pin_memory=pin)
    x.copy_(torch.rand(n, 3, 224, 224, generator=g))
    return x


# ----------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle port of the reference's PyTorch-CPU path on the host cores
# ----------------------------------------------------------------------------------------------------------
def cpu_reference(model_name, K, batch, steps, warmup, budget_s=None):
    """Times `steps` minibatch steps (adil.py:168-188 restated in oracle/adil_oracle.py) on the CPU.  Each step is a
    bounded sample of the workload: `batch` images instead of 100.  Returns (images/s, ms/step, cores, sample)."""
    from dl_attack_on_imagenet_b200.data import build_classifier
    from oracle import adil_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    model = build_classifier(model_name, seed=0, device="cpu")
    net, mean, std = O.split_normalize(model)
    n_img = max(4 * batch, 64)
    x = make_images(n_img, 1, pin=False)
    torch.manual_seed(1234)
    st = O.init_state(3, 224, 224, n_img, K, EPS)
    perm = torch.randperm(n_img)

    def one_step(i):
        idx = perm[(i * batch) % (n_img - batch + 1):][:batch]
        xb = x[idx]
        with torch.no_grad():
            labels = model(xb).argmax(-1)
        xin, _ = O.synth(xb.reshape(batch, P_IMG), st.D2, st.v, idx, mean, std, EPS, O.F_NORMALIZE)
        _, g, _ = O.classifier_grad(net, xin.reshape(batch, 3, 224, 224), labels, 'ce', 50, False, 'sum')
        O.joint_step_(st, g.reshape(batch, P_IMG), idx, 0.01, EPS, std)

    t_w = time.perf_counter()
    for i in range(warmup):
        one_step(i)
    t_w = (time.perf_counter() - t_w) / max(warmup, 1)
    if budget_s is not None and warmup > 0:
        steps = max(1, min(steps, int(budget_s / max(t_w, 1e-3))))
    t0 = time.perf_counter()
    for i in range(steps):
        one_step(warmup + i)
    dt = time.perf_counter() - t0
    sample = ("%d steps x %d images (of the 100-image minibatch), %s, K=%d, N=%d resident images, oracle port of "
              "adil.py:168-188 on torch-CPU" % (steps, batch, model_name, K, n_img))
    return batch * steps / dt, 1e3 * dt / steps, cores, sample, steps


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    value, ms, cores, sample, steps = cpu_reference(args.model, args.atoms, args.ref_batch, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": "ADiL attack images/sec", "value": value, "unit": "images/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "ADiL joint dictionary/code update, random-init %s, %d atoms, l_inf eps=8/255, fp32; "
                               "CPU sample of %d images per step" % (args.model, args.atoms, args.ref_batch)},
        "cpu_baseline": {"value": value, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    from dl_attack_on_imagenet_b200 import ADIL, ops
    from dl_attack_on_imagenet_b200.data import build_classifier

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (our arm) needs a CUDA device: the ADiL kernels have no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner on stdout when the communicator comes up: keep stdout for the ONE JSON line
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.all_reduce(torch.zeros(1, device=dev))
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    torch.backends.cudnn.allow_tf32 = bool(args.tf32)
    torch.backends.cuda.matmul.allow_tf32 = bool(args.tf32)
    torch.backends.cudnn.benchmark = not args.no_cudnn_autotune
    ops.set_impl({"auto": ops.IMPL_AUTO, "fma": ops.IMPL_FMA, "tc": ops.IMPL_TC}[args.kernel_impl])

    B, K, N = args.batch, args.atoms, args.images
    shape = (3, 224, 224)
    model = build_classifier(args.model, seed=0, device=dev)
    for p in model.parameters():
        p.requires_grad_(False)
    x_host = make_images(N, 1 + rank, pin=True)           # this rank's image shard, pinned host memory
    x_dev = x_host.to(dev).view(N, P_IMG)                 # resident copy for the device-timed run
    ADIL.verbose = False
    atk = ADIL(model, eps=EPS, n_atoms=K, batch_size=B, model_name="bench_%d" % rank, step_size=0.01, loss='ce',
               method='gd')
    torch.manual_seed(1234)
    st = atk.begin_fit(N, shape)
    if world > 1:                                         # replicated dictionary: rank 0's draw
        dist.broadcast(st.D, 0)
    g_perm = torch.Generator().manual_seed(7 + rank)
    perm = torch.randperm(N, generator=g_perm)
    n_batches = max(N // B, 1)
    idx_cpu = [perm[(i % n_batches) * B:(i % n_batches) * B + B].contiguous() for i in range(n_batches)]
    idx_dev = [t.to(dev) for t in idx_cpu]
    dD2 = torch.empty_like(st.D2) if world > 1 else None
    flags = ops.SYNTH_NORMALIZE
    mean, std = atk._mean, atk._std
    hbm_peak, peak_src = measured_peaks()

    ev = {k: [] for k in ("synth", "grad", "code", "dict")}

    def timed(name, fn, record):
        if not record:
            return fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = fn()
        b.record()
        ev[name].append((a, b))
        return out

    def step_resident(i, record=False):
        """Device-resident step: images already in HBM, gathered inside the synthesis kernel."""
        idx = idx_dev[i % n_batches]
        idx_h = idx_cpu[i % n_batches]     # the DataLoader's CPU index tensor (adil.py:168): the synthesis / backward
                                           # kernels take it as kernel parameters, like ADIL._fit_step does
        with torch.no_grad():
            labels = model(x_dev[idx].view(-1, *shape)).argmax(-1)               # adil.py:172
        xin, _ = timed("synth", lambda: ops.synth(st.D2, st.v, idx_h, x=x_dev, x_index=idx_h, mean=mean, std=std,
                                                  flags=flags), record)
        loss, g, out = atk._classifier_grad(xin.view(-1, *shape), labels, 'sum')
        g2 = g.view(B, P_IMG)
        if world == 1:
            st.tD += 1
            dvb = timed("grad", lambda: ops.grad_dict_step(st.D2, st.mD, st.sD, g2, st.v, idx_h,
                                                           ops.adamw_params(st.tD, 0.01), std, ops.ATOMS_CLAMP1), record)
        else:
            _, dvb = timed("grad", lambda: ops.grad(g2, st.D2, st.v, idx_h, std, dD2=dD2), record)
            dist.all_reduce(dD2, op=dist.ReduceOp.SUM)                           # the one data-path collective
            st.tD += 1
            timed("dict", lambda: ops.dict_step(st.D2, st.mD, st.sD, dD2, ops.adamw_params(st.tD, 0.01),
                                                ops.ATOMS_CLAMP1), record)
        st.tv += 1
        timed("code", lambda: ops.code_step(st.v, st.mv, st.sv, dvb, idx, ops.adamw_params(st.tv, 0.01),
                                            ops.ROWS_L1BALL, EPS), record)
        return loss

    from dl_attack_on_imagenet_b200.data import HostBatchPrefetcher
    prefetch = HostBatchPrefetcher(x_host, dev)    # public staging helper: pinned gather + H2D on a side stream

    def step_e2e(i, last):
        """End-to-end step through the public API: host gather into pinned memory + H2D (HostBatchPrefetcher: the
        copy of batch i+1 overlaps the kernels of batch i), ADIL.fit_batch (multi-GPU: the same kernels +
        all-reduce), D2H of loss and fooled count."""
        idx = idx_cpu[i % n_batches]
        xb = prefetch.get()
        if world == 1:
            loss, fooled = atk.fit_batch(idx, xb)
        else:
            idd = idx.to(dev, non_blocking=True)
            labels = atk._clean_labels(xb)
            xin, _ = ops.synth(st.D2, st.v, idd, x=xb.view(B, P_IMG), mean=mean, std=std, flags=flags)
            loss, g, out = atk._classifier_grad(xin.view(-1, *shape), labels, 'sum')
            fooled = (out.argmax(-1) != labels).sum()
            _, dvb = ops.grad(g.view(B, P_IMG), st.D2, st.v, idd, std, dD2=dD2)
            dist.all_reduce(dD2, op=dist.ReduceOp.SUM)
            st.tD += 1
            ops.dict_step(st.D2, st.mD, st.sD, dD2, ops.adamw_params(st.tD, 0.01), ops.ATOMS_CLAMP1)
            st.tv += 1
            ops.code_step(st.v, st.mv, st.sv, dvb, idd, ops.adamw_params(st.tv, 0.01), ops.ROWS_L1BALL, EPS)
        prefetch.release()
        if not last:
            prefetch.submit(idx_cpu[(i + 1) % n_batches])   # gathered and copied while the GPU runs step i
        return loss.item(), fooled.item()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    # ---- device-resident timed region -------------------------------------------------------------------
    for i in range(args.warmup):
        step_resident(i)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step_resident(args.warmup + i, record=True)
    e1.record()
    barrier()
    clocks = sampler.finish()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    ms_step = ms_total / args.steps
    value = world * B * args.steps / (ms_total * 1e-3)
    launches_per_step = 4 if world == 1 else 5     # synth, grad(+reduce_partials), [dict], code
    kt = {k: (sum(a.elapsed_time(b) for a, b in v) / len(v)) if v else None for k, v in ev.items()}

    # ---- end-to-end region (host buffers, copies inside) ---------------------------------------------------
    e2e = None
    if not args.no_e2e:
        prefetch.submit(idx_cpu[0])
        for i in range(args.warmup):
            step_e2e(i, last=(i == args.warmup - 1))
        barrier()
        t0 = time.perf_counter()
        e0.record()
        prefetch.submit(idx_cpu[args.warmup % n_batches])   # exactly `steps` H2D copies inside the timed region
        for i in range(args.steps):
            step_e2e(args.warmup + i, last=(i == args.steps - 1))
        e1.record()
        barrier()
        ms_e2e = max_over_ranks(max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3))
        e2e = {"value": world * B * args.steps / (ms_e2e * 1e-3), "unit": "images/s",
               "h2d_bytes_per_step": B * P_IMG * 4 + B * 8, "d2h_bytes_per_step": 4 + 8,
               "ms_per_step": ms_e2e / args.steps}

    # ---- roofline of the dominant ADiL kernel -----------------------------------------------------------------
    if world == 1:
        alg_bytes = 4.0 * P_IMG * (B + 6 * K) + 8.0 * B * K
        kname = "grad_dict_step (dD=g^T v, dv=g D, AdamW(D), clamp fused; adil_grad_dict_step)"
    else:
        alg_bytes = 4.0 * P_IMG * (B + 2 * K) + 8.0 * B * K
        kname = "grad (dD=g^T v, dv=g D; adil_grad) before the NCCL all-reduce"
    synth_bytes = 4.0 * P_IMG * (2 * B + K) + 4.0 * B * K
    roofline = None
    kernels = {}
    if kt["grad"]:
        ach = alg_bytes / (kt["grad"] * 1e-3) / 1e9
        roofline = {"bound": "hbm", "kernel": kname, "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
                    "frac": ach / hbm_peak,
                    "traffic": measured_traffic("adil_grad_dict_step" if world == 1 else "adil_grad"),
                    "alg_bytes": alg_bytes, "kernel_ms": kt["grad"], "peak_source": peak_src}
        kernels["grad"] = {"ms": kt["grad"], "GBps": ach, "frac": ach / hbm_peak}
    if kt["synth"]:
        ach = synth_bytes / (kt["synth"] * 1e-3) / 1e9
        kernels["synth"] = {"ms": kt["synth"], "GBps": ach, "frac": ach / hbm_peak, "alg_bytes": synth_bytes}
    if kt["code"]:
        kernels["code_step"] = {"ms": kt["code"], "alg_bytes": 28.0 * N * K}
    if kt["dict"]:
        ach = 28.0 * P_IMG * K / (kt["dict"] * 1e-3) / 1e9
        kernels["dict_step"] = {"ms": kt["dict"], "GBps": ach, "frac": ach / hbm_peak}

    # ---- CPU baseline (rank 0, N=1 only): oracle port on the host cores, bounded sample ------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            v_cpu, _, cores, sample, _ = cpu_reference(args.model, K, args.ref_batch, 4, 1, budget_s=20.0)
            cpu = {"value": v_cpu, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample}
        except Exception as exc:  # keep the GPU numbers even if the host run fails
            cpu = {"value": None, "unit": "images/s", "cores": os.cpu_count(), "kind": "port", "sample": "failed: %r" % exc}

    if rank == 0:
        line = {
            "metric": "ADiL attack images/sec", "value": value, "unit": "images/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {
                "workload": "BASELINE configs[1]: ADiL joint dictionary/code update ('gd', adil.py:168-188) on "
                            "random-init %s, %d synthetic 3x224x224 images per GPU, %d atoms, minibatch %d per GPU, "
                            "l_inf eps=8/255, AdamW lr 0.01, CE loss" % (args.model, N, K, B),
                "classifier_math": "cuDNN TF32 allowed" if args.tf32 else "strict fp32 (TF32 off)",
                "adil_kernels": "tcgen05 split precision (3xTF32 synthesis, bf16x3 backward), fp32 accumulate; FMA "
                                "fallback for shapes outside B<=128, K<=128 (impl=%s)" % args.kernel_impl,
                "l2": "inputs larger than L2: each step touches >150 MB of ADiL state + GBs of activations",
                "parallelism": "image-sharded x%d, dD SUM all-reduce (NCCL)" % world if world > 1 else "single GPU",
            },
            "e2e": e2e, "gpu_launches": launches_per_step * args.steps, "clocks": clocks, "roofline": roofline,
            "cpu_baseline": cpu, "kernels": kernels,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

"""Turns the raw captures a profiling gpurun call leaves in gpurun_out/ into the tracked summaries under profiles/.

    python scripts/make_profiles.py TAG            # e.g. r01d

expects  gpurun_out/launches_TAG.csv         ncu --metrics gpu__time_duration.sum launch list of bench.py
         gpurun_out/prof_TAG_synth.ncu-rep   ncu --set full capture of synth_kernel  (scripts/kernel_bench.py)
         gpurun_out/prof_TAG_grad.ncu-rep    ncu --set full capture of grad_kernel   (fused step)
writes   profiles/TAG_bench_launches_summary.txt, profiles/TAG_{synth,grad}_kernel_ncu_summary.txt,
         profiles/ncu_traffic.json (read by bench.py for roofline.traffic)
"""
import collections, csv, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
GO, PR = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")


def run(*cmd):
    return subprocess.run(list(cmd), capture_output=True, text=True).stdout


def launches():
    path = os.path.join(GO, f"launches_{tag}.csv")
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 5]
    hdr = next(r for r in rows if "Kernel Name" in r)
    ki, vi, ii = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("ID")
    recs = []
    for r in rows:
        if r is hdr or len(r) <= vi:
            continue
        try:
            recs.append((int(r[ii]), float(r[vi].replace(",", "")) / 1e3, r[ki]))  # ns -> us
        except ValueError:
            pass
    tot = sum(t for _, t, _ in recs)
    out = [f"ncu --metrics gpu__time_duration.sum --clock-control none --csv python bench.py --steps 2 --warmup 3 "
           f"--no-cpu-baseline --no-e2e --no-cudnn-autotune   (kernels inside the replayed CUDA graphs are listed one by one)",
           "(B200; cold-cache serialised per-launch times: compare SHARES, not absolutes)",
           f"{len(recs)} launches, {tot / 1e3:.1f} ms total device time", "",
           "every libadil_b200 launch (ID, us, kernel):"]
    ours = [r for r in recs if "adil::" in r[2]]
    for i, t, k in ours:
        out.append(f"{i:8d} {t:9.2f}  {k[:110]}")
    out += ["", f"libadil_b200 share of device time: {100 * sum(t for _, t, _ in ours) / tot:.3f} %", "",
            "per-kernel totals (top 25):"]
    agg = collections.defaultdict(lambda: [0.0, 0])
    for _, t, k in recs:
        agg[k][0] += t
        agg[k][1] += 1
    for k, (t, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:25]:
        out.append(f"{t / 1e3:10.3f} ms {100 * t / tot:5.1f}% x{n:5d} avg {t / n:9.1f} us  {k[:120]}")
    open(os.path.join(PR, f"{tag}_bench_launches_summary.txt"), "w").write("\n".join(out) + "\n")


def kernel(name, regex, only):
    rep = os.path.join(GO, f"prof_{tag}_{name}.ncu-rep")
    head = (f"ncu --set full --clock-control none --import-source on -k regex:{regex} -s 3 -c 1  python "
            f"scripts/kernel_bench.py --impls auto --only {only} --iters 3   (B=100, K=50, P=150528)\n")
    top = run(sys.executable, os.path.join(ROOT, "scripts", "ncu_top.py"), rep, "14")
    smp = run(sys.executable, os.path.join(ROOT, "scripts", "ncu_samples.py"), rep, "24")
    ops = run(sys.executable, os.path.join(ROOT, "scripts", "ncu_exec.py"), rep)
    raw = run("ncu", "-i", rep, "--page", "raw", "--csv")
    rows = list(csv.reader(io.StringIO(raw)))
    d = dict(zip(rows[0], rows[2]))
    extra = ["", "pipes / shared memory:"]
    for k in rows[0]:
        if any(t in k for t in ("pipe_xu", "pipe_alu_realtime", "pipe_tensor_cycles_active_realtime", "data_pipe_lsu_wavefronts_mem_shared.sum",
                                "data_pipe_tc_wavefronts_mem_shared.sum", "lts__t_bytes.sum ")) and d.get(k):
            extra.append(f"  {k:100s} {d[k]}")
    open(os.path.join(PR, f"{tag}_{name}_kernel_ncu_summary.txt"), "w").write(
        head + top + "\n".join(extra) + "\n\nstall samples / instructions per source line (scripts/ncu_samples.py):\n" + smp +
        "\nexecuted instructions by opcode (scripts/ncu_exec.py):\n" + ops)

    def num(key):
        v, unit = float(d[key].replace(",", "")), rows[1][rows[0].index(key)]
        return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(unit, 1.0)
    return {"kernel": d.get("Kernel Name", "?"), "dram_bytes_read": num("dram__bytes_read.sum"),
            "dram_bytes_write": num("dram__bytes_write.sum"),
            "duration_us_under_ncu": float(d["gpu__time_duration.sum"].replace(",", "")),
            "capture": os.path.basename(rep), "config": "B=100, K=50, P=150528 (scripts/kernel_bench.py)"}


launches()
traffic = {"adil_synth": kernel("synth", "synth_kernel", "synth"),
           "adil_grad_dict_step": kernel("grad", "grad_kernel", "grad_dict_step_contig")}
if os.path.exists(os.path.join(GO, f"prof_{tag}_gradplain.ncu-rep")):
    traffic["adil_grad"] = kernel("gradplain", "grad_kernel", "grad_contig")
json.dump(traffic, open(os.path.join(PR, "ncu_traffic.json"), "w"), indent=1)
print(json.dumps(traffic, indent=1))

"""Summarise an .ncu-rep: headline metrics + top SASS instructions by stall samples (needs ncu on PATH)."""
import csv, subprocess, sys, io
rep = sys.argv[1]; ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); hdr, units = rows[0], rows[1]
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__inst_executed.sum', 'launch__grid_size', 'launch__block_size', 'lts__t_bytes.sum',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_membar_per_issue_active.ratio']
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("=== ", d.get('Kernel Name', '?')[:100])
    for k in keys:
        if k in d: print(f"  {k:86s} {d[k]:>16s} {units[hdr.index(k)]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
starts = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
for si_, i0 in enumerate(starts):
    hdr = rows[i0]; end = starts[si_ + 1] - 1 if si_ + 1 < len(starts) else len(rows)
    si = hdr.index('# Samples'); ie = hdr.index('Instructions Executed')
    cols = {n: hdr.index(n) for n in hdr if n.startswith('stall_') and 'Not Issued' not in n}
    data = []
    for k, r in enumerate(rows[i0 + 1:end]):
        try: data.append((int(r[si]), k, r))
        except Exception: pass
    tot = sum(d[0] for d in data) or 1
    print(f"--- section {si_}: total samples {tot}, {len(data)} SASS instructions")
    agg = {}
    for n, k, r in data:
        for c, i in cols.items():
            try: agg[c] = agg.get(c, 0) + int(r[i])
            except Exception: pass
    print("   stall totals:", {c[6:]: v for c, v in sorted(agg.items(), key=lambda t: -t[1])[:8]})
    for n, k, r in sorted(sorted(data, key=lambda t: -t[0])[:ntop], key=lambda t: t[1]):
        st = " ".join(f"{c[6:]}={r[i]}" for c, i in cols.items() if r[i] not in ('0', ''))
        print(f"{k:6d} {n:6d} {100*n/tot:5.1f}% ex={r[ie]:>8s} {r[1][:64]:64s} {st}")
    break

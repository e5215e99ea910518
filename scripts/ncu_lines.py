"""Per-source-line roll-up of an .ncu-rep (needs -lineinfo): warp instructions executed and stall samples per CUDA line."""
import csv, subprocess, sys, io
rep = sys.argv[1]; ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
sect = -1; hdr = None; data = {}
for r in rows:
    if r and r[0] == "Line No":
        hdr = r; sect += 1; data[sect] = []; continue
    if hdr is None or not r or not r[0].isdigit(): continue
    ie, si = hdr.index("Instructions Executed"), hdr.index("# Samples")
    try: data[sect].append((int(r[ie]), int(r[si]), r[0], r[1]))
    except Exception: pass
for sect, d in data.items():
    tot = sum(x[0] for x in d) or 1; ts = sum(x[1] for x in d) or 1
    if tot < 1000: continue
    print(f"=== section {sect}: {tot} warp-instr, {ts} samples")
    for n, s, ln, src in sorted(d, key=lambda t: -t[0])[:ntop]:
        print(f"{ln:>5s} {n:10d} {100*n/tot:5.1f}%  smp {100*s/ts:5.1f}%  {src.strip()[:120]}")

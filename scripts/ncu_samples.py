"""Per-source-line roll-up of an .ncu-rep sorted by stall SAMPLES (where the time goes), with instruction share."""
import csv, subprocess, io, sys
rep = sys.argv[1]; ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None; data = []
for r in rows:
    if r and r[0] == "Line No": hdr = r; continue
    if hdr is None or not r or not r[0].isdigit(): continue
    try: data.append((int(r[hdr.index('# Samples')]), int(r[hdr.index('Instructions Executed')]), r[0], r[1]))
    except Exception: pass
ts = sum(d[0] for d in data) or 1; ti = sum(d[1] for d in data) or 1
print("samples", ts, "warp-instr", ti)
for s_, i_, ln, src in sorted(data, key=lambda t: -t[0])[:ntop]:
    print(f"{ln:>5s} smp {100*s_/ts:5.1f}%  instr {100*i_/ti:5.1f}%  {src.strip()[:120]}")

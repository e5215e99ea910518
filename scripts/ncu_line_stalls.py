"""Stall reasons per CUDA source line of an .ncu-rep for a line range: python ncu_line_stalls.py rep lo hi [min_samples]"""
import csv, subprocess, io, sys
rep, lo, hi = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]); mins = int(sys.argv[4]) if len(sys.argv) > 4 else 5
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None
for r in rows:
    if r and r[0] == "Line No": hdr = r; cols = {n: hdr.index(n) for n in hdr if n.startswith('stall_') and 'Not Issued' not in n}; continue
    if hdr is None or not r or not r[0].isdigit(): continue
    try: ln = int(r[0]); n = int(r[hdr.index('# Samples')])
    except Exception: continue
    if lo <= ln <= hi and n >= mins:
        st = " ".join(f"{c[6:]}={r[i]}" for c, i in cols.items() if r[i] not in ('0', '') and r[i].isdigit() and int(r[i]) >= 3)
        print(f"{ln:5d} smp {n:5d} ex {r[hdr.index('Instructions Executed')]:>9s}  {r[1].strip()[:60]:60s} | {st}")

"""Executed-instruction histogram by SASS opcode for one kernel section of an .ncu-rep."""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]; sect = int(sys.argv[2]) if len(sys.argv) > 2 else 0
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
starts = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
i0 = starts[sect]; end = starts[sect + 1] - 1 if sect + 1 < len(starts) else len(rows)
hdr = rows[i0]; ie = hdr.index('Instructions Executed')
ops = collections.Counter(); tot = 0
for r in rows[i0 + 1:end]:
    try: n = int(r[ie])
    except Exception: continue
    s = r[1].split()
    op = s[1] if s and s[0].startswith('@') else (s[0] if s else '?')
    ops[op.split('.')[0]] += n; tot += n
print("kernel:", rows[i0 - 1][1][:80] if i0 > 0 else "?", " total warp-instr executed:", tot)
for op, n in ops.most_common(28): print(f"  {op:14s} {n:10d} {100*n/tot:5.1f}%")

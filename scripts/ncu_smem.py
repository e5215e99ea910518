"""Shared-memory wavefronts per SASS instruction of an .ncu-rep (source page): where the excess (bank-conflict) wavefronts are."""
import csv, subprocess, io, sys
rep = sys.argv[1]; ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = next(r for r in rows if r and r[0] == "Address")
iw, ie, ii, ix = hdr.index("L1 Wavefronts Shared"), hdr.index("L1 Wavefronts Shared Excessive"), hdr.index("L1 Wavefronts Shared Ideal"), hdr.index("Instructions Executed")
data = []
for k, r in enumerate(rows[rows.index(hdr) + 1:]):
    try: data.append((int(r[iw]), int(r[ie]), int(r[ii]), int(r[ix]), k, r[1].strip()))
    except Exception: pass
tw = sum(d[0] for d in data); te = sum(d[1] for d in data)
print("total shared wavefronts", tw, "excessive", te)
for w, e, i, x, k, src in sorted(data, key=lambda t: -t[1])[:ntop]:
    print(f"{k:6d} wave {w:9d} excess {e:9d} ideal {i:9d} ex {x:8d}  {src[:90]}")

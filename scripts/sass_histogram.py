"""SASS opcode histogram + resource usage of every kernel in libadil_b200.so (cuobjdump; no GPU needed).

    python scripts/sass_histogram.py > profiles/r02_sass_opcode_histogram.txt

UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st (TMEM), UBLKCP = cp.async.bulk (1-D TMA), UBLKRED = cp.reduce.async.bulk,
LDGSTS = cp.async, LDGMC = multimem.ld_reduce (NVLS), SYNCS = mbarrier ops, UTCBAR = tcgen05.commit, LDL/STL = local
memory (spills: the tcgen05 kernels must have none -- a reload queues behind the stores of the AdamW pass).
"""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "dl_attack_on_imagenet_b200", "libadil_b200.so")
OPS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UBLKCP", "UBLKRED", "UTCBAR", "LDGSTS", "LDGMC", "SYNCS", "HMMA", "FFMA", "LDG",
       "STG", "LDL", "STL"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def short(name):
    name = re.sub(r"adil::\(anonymous namespace\)::", "", name)
    name = re.sub(r"\((anonymous namespace|adil)::[A-Za-z]+Args\)$|\(.*\)$", "", name)
    return name.replace("(int)", "").replace("(bool)", "")[:70]


sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
counts, order, cur = collections.defaultdict(collections.Counter), [], None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); order.append(cur); continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        op = m.group(1)
        counts[cur]["_n"] += 1
        for o in OPS:
            if op == o or op.startswith(o + "."):
                counts[cur][o] += 1
usage = {}
for m in re.finditer(r"Function (\S+?):\s*\n\s*REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", res):
    usage[m.group(1)] = (int(m.group(2)), int(m.group(3)), int(m.group(5)))
names = demangle(order)
print(__doc__.split("\n\n")[0].splitlines()[0])
print("(cuobjdump -sass / -res-usage of %s; nvcc 12.9, -gencode arch=compute_100a,code=sm_100a)" % os.path.relpath(LIB, ROOT))
print("\n".join(__doc__.split("\n\n")[2].splitlines()))
print()
print("%-70s %6s %4s %5s " % ("kernel", "instr", "REG", "STACK") + " ".join("%7s" % o for o in OPS))
for f in order:
    r = usage.get(f, (0, 0, 0))
    print("%-70s %6d %4d %5d " % (short(names[f]), counts[f]["_n"], r[0], r[1]) + " ".join("%7d" % counts[f][o] for o in OPS))

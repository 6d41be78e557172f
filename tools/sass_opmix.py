"""Per-kernel SASS op mix of libmovfem_b200.so (cuobjdump -sass): the instructions that show which hardware paths a kernel uses
and whether it spills -- DFMA/DMUL/DADD (FP64 pipe), DMMA (FP64 tensor-core path), UBLKCP (TMA bulk copy), LDGSTS (cp.async),
SYNCS (mbarrier), BAR (block/named barriers), STL/LDL (local memory = spills), ATOM/RED.

    python tools/sass_opmix.py > profiles/r02_sass_opmix.md
"""
import collections, os, re, subprocess, sys

HERE = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(HERE, "movfem_b200", "libmovfem_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
COLS = ["DFMA", "DMUL", "DADD", "DMMA", "MUFU", "LDS", "STS", "LDG", "STG", "LDGSTS", "UBLKCP", "SYNCS", "BAR", "ATOM", "RED", "STL", "LDL"]
kern, rows = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(.*\)$", "", kern).replace("movfem::", "")
        rows[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and kern:
        op = m.group(1)
        rows[kern]["total"] += 1
        for c in COLS:
            if op == c or op.startswith(c + "."):
                rows[kern][c] += 1
        if op in ("ATOMS", "ATOMG"): rows[kern]["ATOM"] += 1
        if op in ("REDG", "REDS"): rows[kern]["RED"] += 1
print("# SASS op mix per kernel (static instruction counts, `cuobjdump -sass movfem_b200/libmovfem_b200.so`, sm_100a)\n")
print("STL/LDL = local memory (spills); UBLKCP = TMA bulk copy; LDGSTS = cp.async; SYNCS = mbarrier; DMMA = mma.sync.m8n8k4.f64.\n")
print("| kernel | total | " + " | ".join(COLS) + " |")
print("|---|---|" + "|".join(["---"] * len(COLS)) + "|")
for k, c in rows.items():
    if c["total"] < 40: continue
    name = k if len(k) < 90 else k[:87] + "..."
    print(f"| `{name}` | {c['total']} | " + " | ".join(str(c[x]) if c[x] else "" for x in COLS) + " |")

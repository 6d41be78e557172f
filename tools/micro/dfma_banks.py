"""Operand-bank statistics of the DFMA instructions in a binary (for dfma_issue_bench / tile_bench / the library).

    python tools/micro/dfma_banks.py tools/micro/dfma_issue_bench [function-substring]

Hypothesis to test against the measured cycles per DFMA: the register file delivers one 64-bit operand per bank and
cycle, bank = (register / 2) % 2; operands flagged .reuse come from the reuse cache; a DFMA whose NEW operands (at least
two) all sit in one bank needs an extra cycle.  Prints per function: DFMAs, how many have all new operands in one bank."""
import re
import subprocess
import sys

sass = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
want = sys.argv[2] if len(sys.argv) > 2 else ""
for fn in re.split(r"Function : ", sass)[1:]:
    name = fn.split("\n")[0]
    if want not in name:
        continue
    n = conf = 0
    for m in re.finditer(r"DFMA R(\d+), (-?)R(\d+)(\.reuse)?, (-?)R(\d+)(\.reuse)?, (-?)R(\d+)(\.reuse)?", fn):
        _, _, a, ar, _, b, br, _, c, cr = m.groups()
        new = [int(x) for x, r in ((a, ar), (b, br), (c, cr)) if not r]
        n += 1
        if len(new) >= 2 and len({(x // 2) % 2 for x in new}) == 1:
            conf += 1
    if n:
        print(f"{name[:90]:90s} DFMA {n:5d}  all-new-operands-in-one-bank {conf:5d} ({100.0 * conf / n:.0f} %)")

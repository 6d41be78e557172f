"""Writes tests/golden/ref_module_uses.json: which module every source file of the reference defines and which modules it
`use`s (read from /root/reference in place).  tests/test_c_harness.py sorts this graph together with the Fortran shim."""
import glob, json, os, re

SRC = "/root/reference/MoVFEM_3DMT/src"
out = {}
for f in sorted(glob.glob(os.path.join(SRC, "*.f90"))):
    text = open(f, errors="replace").read()
    code = "\n".join(line.split("!")[0] for line in text.splitlines())
    units = re.findall(r"^\s*(module|program)\s+(\w+)", code, flags=re.I | re.M)
    units = [(k.lower(), n.lower()) for k, n in units if n.lower() != "procedure"]
    uses = sorted({u.lower() for u in re.findall(r"^\s*use\s*(?:,\s*intrinsic\s*::)?\s*(\w+)", code, flags=re.I | re.M)})
    out[os.path.basename(f)] = {"defines": [n for _, n in units], "uses": uses}
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_module_uses.json"), "w"), indent=1, sort_keys=True)
print(json.dumps(out, indent=1))

"""Per-phase timing of geometry_kernel (profiling aid): cold assemblies with MOVFEM_PHASE_MASK (bit 0: interpolation +
Jacobian/tensor phases, bit 1: RHS phase); prints the geometry and contraction times of each mask."""
import os, sys, subprocess, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
cfg = sys.argv[1] if len(sys.argv) > 1 else "2"
dirich = sys.argv[2] if len(sys.argv) > 2 else "None"
code = r'''
import sys, os, numpy as np, torch
sys.path.insert(0, %r)
from movfem_b200 import mesh, host, abi
m = mesh.config(%s, dirichlet=%s)
asm = host.Assembly(m)
om, sg = m.omega(1), m.sigma_for(1)
d = torch.from_numpy(sg.view(np.float64).reshape(-1).copy()).cuda()
ts = []
for it in range(6):
    asm.reset_cache(); asm.assemble_device(1, om, d.data_ptr(), abi.MODE_T2)
    try: asm.device_result()
    except Exception as e: pass
    ts.append((asm.stats()["ms_geometry"], asm.stats()["ms_contract"]))
print("geometry %.4f contraction %.4f" % min(ts[2:]))
''' % (ROOT, cfg, dirich)
for mask, name in ((0, "node staging only"), (1, "staging + B1 + B2"), (2, "staging + RHS"), (3, "all")):
    env = dict(os.environ, MOVFEM_PHASE_MASK=str(mask))
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    print(f"mask {mask:2d} {name:26s} {out.stdout.strip()} ms", out.stderr.strip()[-200:] if out.returncode else "")

"""Run N cold assemblies of one config (profiling driver): python tools/run_one.py <config> <dirichlet|None> [n]"""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from movfem_b200 import mesh, host, abi
cfg = int(sys.argv[1]); dirich = None if len(sys.argv) < 3 or sys.argv[2] == "None" else int(sys.argv[2]); n = int(sys.argv[3]) if len(sys.argv) > 3 else 3
m = mesh.config(cfg, dirichlet=dirich)
asm = host.Assembly(m)
om, sg = m.omega(1), m.sigma_for(1)
d = torch.from_numpy(sg.view(np.float64).reshape(-1).copy()).cuda()
for it in range(n):
    asm.reset_cache(); asm.assemble_device(1, om, d.data_ptr(), abi.MODE_T2); asm.device_result()
print(asm.stats())

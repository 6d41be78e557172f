"""Small assemblies of every element type / boundary mode for compute-sanitizer (memcheck, racecheck, synccheck)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from movfem_b200 import mesh, host, abi
for mn, dirichlet, sch, inimod in ((8, 0, 1, 1), (20, 0, 0, 1), (27, 1, 1, 3), (20, 1, 1, 2)):
    m = mesh.build_model(f"san_mn{mn}", 6, 5, mn, 1000., 1100., 900., 2, 2, 1, dirichlet=dirichlet, gpml_sch=sch, freqs=(0.5, 3.0),
                         sigma_fn=mesh._layered((1500., 1500., 500., 1500.)), topo_amp=50.0)
    m.bd_inimod = inimod; m.bd_lsigma, m.bd_ldz = (0.01, 0.1), (1.5,)
    asm = host.Assembly(m)
    for f in (1, 2, 2):
        r = asm.global_vfem(f, m.omega(f), m.sigma_for(f))
    print(mn, dirichlet, sch, inimod, r[4], float(np.abs(r[2][: r[4]]).max()))
    asm.close()

"""Hash of the assembly results of the library selected by MOVFEM_B200_LIB (A/B builds must be bit-identical to the
default: warp counts, cache policies, tile shapes do not change any summation order).  Prints one line per case:
sha256 over gne, irn, jcn, a (T1 and T2) and rhs of two frequencies.  Needs a GPU."""
import hashlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from movfem_b200 import abi, host, mesh

for mn, dirichlet, sch in ((8, 0, 0), (20, 0, 0), (20, 1, 1), (27, 0, 0), (27, 0, 1)):
    m = mesh.build_model(f"abhash_mn{mn}", 7, 6, mn, 1000., 1100., 900., 2, 2, 1, dirichlet=dirichlet, gpml_sch=sch, freqs=(0.5, 3.0),
                         sigma_fn=mesh._layered((1500., 1500., 500., 1500.)), topo_amp=40.0)
    asm = host.Assembly(m)
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(asm.gne()).tobytes())
    for ifreq in (1, 2):
        for mode in (abi.MODE_T1, abi.MODE_T2):
            irn, jcn, a, rhs, nz = asm.global_vfem(ifreq, m.omega(ifreq), m.sigma_for(ifreq), mode=mode)
            for x in (irn[:nz], jcn[:nz], a[:nz], rhs):
                h.update(np.ascontiguousarray(x).tobytes())
    asm.close()
    print(f"mn{mn}_d{dirichlet}_s{sch} {h.hexdigest()[:16]}")

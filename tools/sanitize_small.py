"""Small assemblies of every element type / boundary mode for compute-sanitizer (memcheck, racecheck, synccheck)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from movfem_b200 import mesh, host, abi
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden"))
import make_reference_vectors as mrv
for mn, dirichlet, sch, inimod in ((8, 0, 1, 1), (20, 0, 0, 1), (27, 1, 1, 3), (20, 1, 1, 2), (27, 0, 0, 1), (8, 0, 0, 1)):
    m = mesh.build_model(f"san_mn{mn}", 6, 5, mn, 1000., 1100., 900., 2, 2, 1, dirichlet=dirichlet, gpml_sch=sch, freqs=(0.5, 3.0),
                         sigma_fn=mesh._layered((1500., 1500., 500., 1500.)), topo_amp=50.0)
    m.bd_inimod = inimod; m.bd_lsigma, m.bd_ldz = (0.01, 0.1), (1.5,)
    asm = host.Assembly(m)
    for f in (1, 2, 2):       # cold pass, flags of element (1,1,1) change (scheme 0: one more full pass), cached RHS-only pass
        r = asm.global_vfem(f, m.omega(f), m.sigma_for(f))
    print(mn, dirichlet, sch, inimod, r[4], float(np.abs(r[2][: r[4]]).max()))
    asm.close()
# linear elements, many batches per CTA of fused12_kernel (MOVFEM_TEST_FUSED_GRID=2): its mbarrier pipeline; both sigma variants
os.environ["MOVFEM_TEST_FUSED_GRID"] = "2"
for aniso in (False, True):
    m = mesh.build_model("san_fused", 14, 9, 8, 1000., 1100., 900., 2, 2, 1, dirichlet=0, gpml_sch=1, freqs=(0.5, 3.0),
                         sigma_fn=mesh._layered((1500., 1500., 500., 1500.)), topo_amp=50.0)
    if aniso: m.sigma_re[:, 3] *= 1.5
    asm = host.Assembly(m)
    for f in (1, 2, 2):
        r = asm.global_vfem(f, m.omega(f), m.sigma_for(f))
    print("fused12 pipeline", aniso, m.ne, r[4], float(np.abs(r[2][: r[4]]).max()))
    asm.close()
del os.environ["MOVFEM_TEST_FUSED_GRID"]
# geomodel -> grid nodes (movfem_geo_innermodel)
for name in sorted(mrv.GEO_CASES):
    m, n_air, inp = mrv.geo_case(name)
    S, M, ms = host.innermodel_gqg(m, n_air, m.omega(1), inp["xm"], inp["ym"], inp["zm"], inp["ijsigma"], inp["sigma"], inp["ijmu"], inp["mu"])
    print(name, S.shape, float(np.abs(S).max()))

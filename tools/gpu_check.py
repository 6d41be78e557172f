"""Quick GPU-vs-oracle check used during development (run under gpurun)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from movfem_b200 import mesh, host, abi
from oracle.oracle import Oracle
from parity_util import compare_assembly, rel_err

def check(model, label, elems=(1, 2)):
    t0 = time.time()
    o = Oracle(model)
    asm = host.Assembly(model)
    g1, g2 = asm.gne(), o.gne()
    print(f"== {label}: ne={model.ne} nne={asm.nne}/{o.nne} nnze={asm.nnze}/{o.nnze} nzu={asm.nz_upper}/{o.nz_upper} gne_equal={np.array_equal(g1,g2)}")
    r = compare_assembly(asm, o, model)
    for k, v in r.items(): print(f"   {k}: {v}")
    # element taps
    omega = model.omega(1); sigma = model.sigma_for(1)
    for ide in elems:
        if ide > model.ne: continue
        pml = o.effective_pml(ide) if not model.dirichlet else (0, 0, 0)
        if ide == 1: pml = (0,0,0)
        e = o.element(ide, omega, sigma, pml)
        K, M, b = asm.debug_element(ide)
        Ko = e["A"].real; Mo = e["A"].imag / float(np.float32(omega))
        Ko = np.tril(Ko) + np.tril(Ko, -1).T; Mo = np.tril(Mo) + np.tril(Mo, -1).T
        print(f"   elem {ide} pml={tuple(pml)}: K rel {rel_err(K, Ko):.2e} M rel {rel_err(M, Mo):.2e} b rel {rel_err(b, e['b']):.2e}")
    print(f"   stats {asm.stats()}  ({time.time()-t0:.1f}s)")
    asm.close()

if __name__ == "__main__":
    print(host.lib().movfem_version().decode())
    which = sys.argv[1:] or ["s"]
    if "s" in which:
        for mn in (8, 20, 27):
            for dirich in (0, 1):
                m = mesh.build_model(f"small_mn{mn}_d{dirich}", 6, 5, mn, 1000., 1100., 900., 2, 2, 1, dirichlet=dirich, gpml_sch=1, freqs=(0.5,),
                                     sigma_fn=mesh._layered((1500., 1500., 500., 1500.)), topo_amp=50.0)
                check(m, m.name, elems=(1, 2, 7, m.ne))
        m = mesh.build_model("small_mn20_fang", 6, 5, 20, 1000., 1100., 900., 2, 2, 1, dirichlet=0, gpml_sch=0, freqs=(0.5,))
        check(m, m.name, elems=(1, 2, m.ne))
        m = mesh.config(3, scale=0.3)
        check(m, m.name, elems=(1, 2, m.ne))
    if "c1" in which:
        for d in (0, 1):
            m = mesh.config(1, dirichlet=d)
            check(m, m.name + f"_d{d}", elems=(1, 2, 100, m.ne))

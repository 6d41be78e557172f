import sys, os, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from movfem_b200 import host, mesh, abi
from oracle.oracle import Oracle
model = mesh.build_model("smoke_mn20", 6, 5, 20, 1000., 1100., 900., 2, 2, 1, dirichlet=0, gpml_sch=0, freqs=(0.5,),
                         sigma_fn=mesh._layered((1500., 1500., 500., 1500.)), topo_amp=40.0)
asm = host.Assembly(model, device=0); o = Oracle(model)
om, sg = model.omega(1), model.sigma_for(1)
res = o.assemble(om, sg)
irn, jcn, a, rhs, nz = asm.global_vfem(1, om, sg, mode=abi.MODE_T2)
ka = irn[:nz].astype(np.int64) * (asm.nne + 1) + jcn[:nz]
kb = res["irn"].astype(np.int64) * (asm.nne + 1) + res["jcn"]
only = np.setdiff1d(ka, kb)
gne = o.gne()
scale = np.abs(res["a"]).max()
irn1, jcn1, a1, _, nz1 = asm.global_vfem(1, om, sg, mode=abi.MODE_T1)
ia, ja = o.pattern()
for k in only:
    r, c = int(k // (asm.nne + 1)), int(k % (asm.nne + 1))
    i = np.flatnonzero(ka == k)[0]
    print("extra entry", r, c, "graft T2 value", a[i], "scale", scale)
    j = np.flatnonzero((irn1[:nz1] == r) & (jcn1[:nz1] == c))[0]
    jo = np.flatnonzero((ia == c) & (ja == r))[0]
    print("  T1 graft", a1[j], " T1 oracle", res["a_t1"][jo])
    els = [e for e in range(model.ne) if (gne[e] == r).any() and (gne[e] == c).any()]
    for e in els:
        ir, ic = int(np.flatnonzero(gne[e] == r)[0]), int(np.flatnonzero(gne[e] == c)[0])
        pml = o.effective_pml(e + 1)
        eo = o.element(e + 1, om, sg, pml=pml)
        K, M, b = asm.debug_element(e + 1)
        nz_ = model.g_nz - 1; ny_ = model.g_ny - 1
        ie, je, ke = e // (ny_ * nz_) + 1, (e // nz_) % ny_ + 1, e % nz_ + 1
        print("  element", e + 1, (ie, je, ke), "pml", tuple(pml), "local", ir + 1, ic + 1, "oracle A", eo["A"][max(ir, ic), min(ir, ic)], "graft K,M", K[max(ir, ic), min(ir, ic)], M[max(ir, ic), min(ir, ic)])

"""Regenerates tests/golden/*.npz.

The reference (Fortran, no compiler in this image, no shipped outputs) cannot produce vectors, so
two kinds of fixtures are kept:
  * pins_appB.json -- the element-matrix known answers of SURVEY.md App. B item 4, which the survey
    obtained by executing a mechanical translation of the reference's own shape-function tables;
  * small_*.npz    -- outputs of the CPU oracle (oracle/movfem_oracle.cpp) on tiny meshes, one per
    element type / boundary mode, frozen so that oracle refactors cannot silently change results.
Run:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from movfem_b200 import mesh  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402

CASES = {
    "small_mn8_gpml_zhou": dict(mn=8, dirichlet=0, gpml_sch=1),
    "small_mn8_dirichlet": dict(mn=8, dirichlet=1, gpml_sch=1),
    "small_mn20_gpml_fang": dict(mn=20, dirichlet=0, gpml_sch=0),
    "small_mn27_gpml_zhou": dict(mn=27, dirichlet=0, gpml_sch=1),
}


def make_model(mn, dirichlet, gpml_sch):
    if mn == 8:     # 5x5x6 elements, nextd=2 (upper-side GPML only)
        return mesh.build_model("golden", 5, 5, mn, 1000., 1100., 900., 2, 1, 1, dirichlet=dirichlet, gpml_sch=gpml_sch, freqs=(0.5, 2.0),
                                sigma_fn=mesh._layered((1500., 1500., 0., 900.)), topo_amp=40.0)
    # 3x3x4 elements, nextd=1: also exercises the lower-side (-1) GPML branch of boundary_conds.f90:76-81
    return mesh.build_model("golden", 3, 3, mn, 1000., 1100., 900., 1, 1, 1, dirichlet=dirichlet, gpml_sch=gpml_sch, freqs=(0.5, 2.0),
                            sigma_fn=mesh._layered((600., 600., 0., 900.)), topo_amp=40.0)


def main():
    for name, kw in CASES.items():
        m = make_model(**kw)
        o = Oracle(m)
        out = dict(gne=o.gne(), nne=o.nne, nnze=o.nnze)
        for ifreq in (1, 2):
            r = o.assemble(m.omega(ifreq), m.sigma_for(ifreq), faithful=True)
            out[f"irn{ifreq}"] = r["irn"]; out[f"jcn{ifreq}"] = r["jcn"]; out[f"a{ifreq}"] = r["a"]; out[f"rhs{ifreq}"] = r["rhs"]
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, "nne", o.nne, "nnze", o.nnze, "nz", out["a1"].size)


if __name__ == "__main__":
    main()

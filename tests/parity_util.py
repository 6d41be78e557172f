"""Shared parity helpers: compare the CUDA path (through the C ABI) with the CPU oracle."""
import numpy as np

from movfem_b200 import abi


def rel_err(x, y):
    """normwise relative error ||x-y||_inf / ||y||_inf (SURVEY section 7 'hard parts')."""
    d = np.abs(np.asarray(x) - np.asarray(y)).max() if np.size(x) else 0.0
    n = np.abs(np.asarray(y)).max() if np.size(y) else 0.0
    return d / n if n else d


def oracle_upper_t1(o, res):
    """oracle T1 values re-keyed to the graft's delivery order: upper (r<=c) carries A_lower(c,r)."""
    ia, ja = o.pattern()
    low = ia >= ja
    r, c, v = ja[low], ia[low], res["a_t1"][low]          # transposed: row=small id, col=large id
    order = np.lexsort((c, r))
    return r[order], c[order], v[order]


def compare_assembly(asm, o, model, ifreq=1, check_t1=True, faithful=False, oracle_res=None):
    """Runs both sides at frequency ifreq and returns a dict of error figures; asserts nothing."""
    omega = model.omega(ifreq)
    sigma = model.sigma_for(ifreq)
    res = oracle_res if oracle_res is not None else o.assemble(omega, sigma, faithful=faithful)
    out = {}
    # structural pattern
    irn_s, jcn_s = asm.pattern()
    r1, c1, v1 = oracle_upper_t1(o, res)
    out["pattern_equal"] = bool(irn_s.size == r1.size and np.array_equal(irn_s, r1) and np.array_equal(jcn_s, c1))
    out["nne"] = (asm.nne, o.nne)
    out["nnze"] = (asm.nnze, o.nnze)
    out["nz_upper"] = (asm.nz_upper, o.nz_upper)
    if check_t1:
        irn, jcn, a, rhs, nz = asm.global_vfem(ifreq, omega, sigma, mode=abi.MODE_T1)
        out["t1_nz"] = (nz, int(r1.size))
        out["t1_idx_equal"] = bool(nz == r1.size and np.array_equal(irn[:nz], r1) and np.array_equal(jcn[:nz], c1))
        ok = nz == r1.size
        out["t1_rel"] = rel_err(a[:nz], v1) if ok else np.inf
        out["t1_rel_re"] = rel_err(a[:nz].real, v1.real) if ok else np.inf
        out["t1_rel_im"] = rel_err(a[:nz].imag, v1.imag) if ok else np.inf
        out["rhs_rel"] = rel_err(rhs, res["rhs"])
    irn, jcn, a, rhs, nz = asm.global_vfem(ifreq, omega, sigma, mode=abi.MODE_T2)
    out["t2_nz"] = (nz, res["nz"])
    same_idx = nz == res["nz"] and np.array_equal(irn[:nz], res["irn"]) and np.array_equal(jcn[:nz], res["jcn"])
    out["t2_idx_equal"] = bool(same_idx)
    if same_idx:
        av, bv = a[:nz], res["a"]
        neq = np.count_nonzero((av.real != bv.real) | (av.imag != bv.imag))
        out["t2_values_differ"] = int(neq)            # float32 roundings that flipped
        out["t2_rel"] = rel_err(av, bv)
    else:
        # key-wise comparison: entries present on one side only must be noise
        ka = irn[:nz].astype(np.int64) * (asm.nne + 1) + jcn[:nz]
        kb = res["irn"].astype(np.int64) * (asm.nne + 1) + res["jcn"]
        common, ia_, ib_ = np.intersect1d(ka, kb, return_indices=True)
        out["t2_common"] = int(common.size)
        out["t2_rel"] = rel_err(a[:nz][ia_], res["a"][ib_])
        only_a = np.setdiff1d(np.arange(nz), ia_)
        only_b = np.setdiff1d(np.arange(res["nz"]), ib_)
        scale = np.abs(res["a"]).max()
        out["t2_only_graft"] = (int(only_a.size), float(np.abs(a[:nz][only_a]).max() / scale) if only_a.size else 0.0)
        out["t2_only_oracle"] = (int(only_b.size), float(np.abs(res["a"][only_b]).max() / scale) if only_b.size else 0.0)
    out["rhs_rel_t2"] = rel_err(rhs, res["rhs"])
    return out

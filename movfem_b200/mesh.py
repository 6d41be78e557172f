"""Synthetic mesh / model generator: the arrays geometry.f90 hands to the assembly path.

geometry.f90 (with toms660 QSHEP2D) cannot be run here (no Fortran), and it is outside the
hot path: the Fortran host keeps it.  This module only *mimics its outputs* so that the
assembly can be exercised on the five BASELINE.json configurations:

* node lines ``g_xp(nnx)``, ``g_yp(nny)`` are tensor-product; every node has its own
  ``g_zp`` (z fastest: ``id=(ii-1)*g_nyz+(jj-1)*g_nnz+kk``, geometry.f90:517-521)
* ``nextd`` extension cells per side with widths ``1.3*i*dx`` growing outward
  (geometry.f90:268-317); z groups are uniform per column (geometry.f90:526-583)
* ``g_sigma(6,npt)`` complex128, ``g_mu(6,npt)`` float64, packing 11,12,13,22,23,33
  (geometry.f90:1055-1060); diagonal sigma gets ``+ i*f32(eps*omega_1)``; node planes above
  the surface plane are air (geometry.f90:934-947)
* ``update_sigma`` (geometry.f90:144-153) indexes ``g_sigma(i,j)`` with ``i<=g_npt, j<=6`` on a
  ``(6,npt)`` array, so only the first ``npt+30`` linear entries are refreshed
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from .abi import MovfemDesc

PI = 3.1415926535897932384626433          # geometry.f90:25
EPS0 = 8.854187817e-12                    # geometry.f90:25
MU0 = 4.0 * PI * 1.0e-7                   # geometry.f90:26

ELEMENT_TYPES = {8: (12, 2), 20: (36, 3), 27: (54, 3)}   # mn -> (me, nord)


def f32(x):
    """Fortran ``cmplx(x,y)`` without KIND rounds through default real."""
    return float(np.float32(x))


@dataclass
class Model:
    """Everything the C-ABI descriptor needs, as NumPy arrays (kept alive here)."""
    name: str
    g_nx: int
    g_ny: int
    g_nz: int
    mn: int
    nextd: int
    nzl_top: int
    dirichlet: int
    gpml_sch: int
    a0: float
    b0: float
    nn: float
    g_xp: np.ndarray
    g_yp: np.ndarray
    g_zp: np.ndarray
    g_mu: np.ndarray            # (npt, 6) C-order == Fortran (6, npt)
    sigma_re: np.ndarray        # (npt, 6) real part of g_sigma
    sigma_im_mask: np.ndarray   # (npt, 6) bool: entries that carry i*f32(eps*omega)
    freqs: np.ndarray
    bd_inimod: int = 1
    g_ztop: float = 0.0                      # geometry.f90:394: lowest point of the topography interface
    bd_hsigma: float = 0.01                  # PARAM.INP: half-space conductivity (boundary model 2)
    bd_lsigma: tuple = (0.01,)               # PARAM.INP: layer conductivities (boundary model 3)
    bd_ldz: tuple = ()                       # layer thicknesses as handed to bd_setmodel (nl-1 values)
    ie_lo: int = 0
    ie_hi: int = 0
    _sigma_state: np.ndarray | None = field(default=None, repr=False)

    @property
    def me(self):
        return ELEMENT_TYPES[self.mn][0]

    @property
    def nord(self):
        return ELEMENT_TYPES[self.mn][1]

    @property
    def ne(self):
        return (self.g_nx - 1) * (self.g_ny - 1) * (self.g_nz - 1)

    @property
    def npt(self):
        return self.g_zp.size

    def desc(self) -> MovfemDesc:
        d = MovfemDesc()
        d.g_nx, d.g_ny, d.g_nz = self.g_nx, self.g_ny, self.g_nz
        d.nord, d.mn, d.me = self.nord, self.mn, self.me
        d.nextd, d.nzl_top = self.nextd, self.nzl_top
        d.dirichlet, d.bd_inimod, d.gpml_sch = self.dirichlet, self.bd_inimod, self.gpml_sch
        d.sym, d.ndir, d.pe_sch = 1, 2, 1          # MoVFEM_3DMT.f90:300,369
        d.a0, d.b0, d.nn = self.a0, self.b0, self.nn
        d.g_xp = self.g_xp.ctypes.data_as(C.c_void_p)
        d.g_yp = self.g_yp.ctypes.data_as(C.c_void_p)
        d.g_zp = self.g_zp.ctypes.data_as(C.c_void_p)
        d.g_mu = self.g_mu.ctypes.data_as(C.c_void_p)
        d.ie_lo, d.ie_hi = self.ie_lo, self.ie_hi
        d.g_ztop, d.bd_hsigma, d.bd_nl = self.g_ztop, self.bd_hsigma, len(self.bd_lsigma)
        for i, v in enumerate(self.bd_lsigma):
            d.bd_lsigma[i] = v
        for i, v in enumerate(self.bd_ldz):
            d.bd_ldz[i] = v
        return d

    def omega(self, ifreq: int) -> float:
        """geometry.f90:137-140 update_omega (ifreq is 1-based)."""
        return 2.0 * PI * float(self.freqs[ifreq - 1])

    def sigma_initial(self) -> np.ndarray:
        """g_sigma as grid_3d leaves it: imaginary parts built with the FIRST frequency."""
        w1 = self.omega(1)
        s = self.sigma_re.astype(np.complex128)
        s[self.sigma_im_mask] += 1j * f32(EPS0 * w1)
        return s

    def sigma_for(self, ifreq: int) -> np.ndarray:
        """g_sigma seen by the assembly at frequency ``ifreq`` after the reference's sequential
        ``update_sigma`` calls 1..ifreq (geometry.f90:144-153).  Each call rewrites
        the imaginary part of only the first ``npt+30`` linear entries that have one."""
        s = self.sigma_initial()
        flat = s.reshape(-1)
        nlin = min(flat.size, self.npt + 30)
        w = self.omega(ifreq)
        head = flat[:nlin]
        m = head.imag != 0.0
        head[m] = head[m].real + 1j * f32(EPS0 * w)
        return s


def _lines(n_inner_lines: int, d: float, nextd: int) -> np.ndarray:
    """geometry.f90:268-317: nextd cells of width 1.3*i*d (i = nextd..1 inward / 1..nextd outward)."""
    ext = [1.3 * i * d for i in range(1, nextd + 1)]
    widths = ext[::-1] + [d] * (n_inner_lines - 1) + ext
    x = np.concatenate([[0.0], np.cumsum(widths)])
    return x - 0.5 * (x[0] + x[-1])          # origin at the centre of the extended domain


def _refine(lines: np.ndarray, nord: int) -> np.ndarray:
    if nord == 2:
        return lines.copy()
    out = np.empty(2 * lines.size - 1)
    out[0::2] = lines
    out[1::2] = 0.5 * (lines[:-1] + lines[1:])   # mid nodes are arithmetic mid-points (geometry.f90:715-768)
    return out


def build_model(name, nx, ny, mn, dx, dy, dz, nextd, n_earth, n_air, *, dirichlet=0, gpml_sch=1,
                a0=1.0, b0=1.0, nn=2.0, freqs=(0.1,), sigma_fn=None, topo_amp=0.0, seed=None,
                aniso=False) -> Model:
    """nx, ny: ELEMENTS per horizontal axis (including 2*nextd extension cells);
    vertical layers = nextd (bottom ext) + n_earth + n_air + nextd (top ext)."""
    me, nord = ELEMENT_TYPES[mn]
    nz = nextd + n_earth + n_air + nextd
    xl = _lines(nx - 2 * nextd + 1, dx, nextd)
    yl = _lines(ny - 2 * nextd + 1, dy, nextd)
    ext_h = 1.3 * dz * sum(range(1, nextd + 1)) / nextd        # uniform extension layers
    zl = np.concatenate([[0.0], np.cumsum([ext_h] * nextd + [dz] * n_earth + [dz] * n_air + [ext_h] * nextd)])
    g_xp, g_yp, zcol = _refine(xl, nord), _refine(yl, nord), _refine(zl, nord)
    nnx, nny, nnz = g_xp.size, g_yp.size, zcol.size
    ksurf = (nextd + n_earth) * (nord - 1)                       # 0-based node plane of the surface
    z_surface = zcol[ksurf]
    X, Y = np.meshgrid(g_xp, g_yp, indexing="ij")
    Z = np.broadcast_to(zcol, (nnx, nny, nnz)).copy()
    if topo_amp:
        lx, ly = g_xp[-1] - g_xp[0], g_yp[-1] - g_yp[0]
        taper = np.where(zcol <= z_surface, zcol / z_surface, (zcol[-1] - zcol) / (zcol[-1] - z_surface))
        Z += (topo_amp * np.sin(2 * np.pi * X / lx) * np.cos(2 * np.pi * Y / ly))[:, :, None] * taper[None, None, :]
    g_zp = np.ascontiguousarray(Z.reshape(-1))
    npt = g_zp.size

    depth = Z[:, :, ksurf][:, :, None] - Z                      # >= 0 in the earth
    air = np.zeros((nnx, nny, nnz), bool)
    air[:, :, ksurf + 1:] = True                                 # planes ABOVE the surface plane (geometry.f90:934-947)
    sig = np.zeros((nnx, nny, nnz, 6))
    iso = np.full((nnx, nny, nnz), 0.01) if sigma_fn is None else sigma_fn(X[:, :, None] + 0 * Z, Y[:, :, None] + 0 * Z, depth)
    if aniso:
        rng = np.random.default_rng(seed)
        ang = rng.uniform(0, 2 * np.pi, size=(nnx, nny, nnz, 3))
        s = np.exp(rng.uniform(np.log(0.5), np.log(2.0), size=(nnx, nny, nnz)))
        ca, sa = np.cos(ang), np.sin(ang)
        # R = Rz(a) Ry(b) Rx(c)
        Rz = np.zeros((nnx, nny, nnz, 3, 3)); Ry = np.zeros_like(Rz); Rx = np.zeros_like(Rz)
        Rz[..., 0, 0] = ca[..., 0]; Rz[..., 0, 1] = -sa[..., 0]; Rz[..., 1, 0] = sa[..., 0]; Rz[..., 1, 1] = ca[..., 0]; Rz[..., 2, 2] = 1
        Ry[..., 0, 0] = ca[..., 1]; Ry[..., 0, 2] = sa[..., 1]; Ry[..., 2, 0] = -sa[..., 1]; Ry[..., 2, 2] = ca[..., 1]; Ry[..., 1, 1] = 1
        Rx[..., 1, 1] = ca[..., 2]; Rx[..., 1, 2] = -sa[..., 2]; Rx[..., 2, 1] = sa[..., 2]; Rx[..., 2, 2] = ca[..., 2]; Rx[..., 0, 0] = 1
        R = Rz @ Ry @ Rx
        D = np.zeros((3, 3)); D[0, 0], D[1, 1], D[2, 2] = 1.0, 0.1, 0.01
        T = (R @ D @ np.swapaxes(R, -1, -2)) * (s * iso)[..., None, None]
        for k, (p, q) in enumerate([(0, 0), (0, 1), (0, 2), (1, 1), (1, 2), (2, 2)]):
            sig[..., k] = T[..., p, q]
    else:
        for k in (0, 3, 5):
            sig[..., k] = iso
    sig[air] = 0.0
    sigma_re = np.ascontiguousarray(sig.reshape(npt, 6))
    mask = np.zeros((npt, 6), bool)
    mask[:, [0, 3, 5]] = True                                    # diagonals carry i*f32(eps*omega)
    g_mu = np.zeros((npt, 6))
    g_mu[:, [0, 3, 5]] = MU0
    g_ztop = float(Z[:, :, ksurf].min())
    return Model(name=name, g_nx=nx + 1, g_ny=ny + 1, g_nz=nz + 1, mn=mn, nextd=nextd, nzl_top=nextd, g_ztop=g_ztop,
                 dirichlet=dirichlet, gpml_sch=gpml_sch, a0=a0, b0=b0, nn=nn,
                 g_xp=np.ascontiguousarray(g_xp), g_yp=np.ascontiguousarray(g_yp), g_zp=g_zp, g_mu=g_mu,
                 sigma_re=sigma_re, sigma_im_mask=mask, freqs=np.asarray(freqs, float))


def _layered(block=None):
    """0.01 / 0.1 / 0.001 S/m at depth 0-5 / 5-15 / >15 km (+ optional 1 S/m block)."""
    def fn(x, y, depth):
        s = np.where(depth < 5000.0, 0.01, np.where(depth < 15000.0, 0.1, 0.001))
        if block is not None:
            hx, hy, d0, d1 = block
            s = np.where((np.abs(x) <= hx) & (np.abs(y) <= hy) & (depth >= d0) & (depth <= d1), 1.0, s)
        return s
    return fn


def config(n: int, *, scale: float = 1.0, dirichlet: int | None = None) -> Model:
    """The five BASELINE.json configurations.  ``scale`` < 1 shrinks the element
    counts (never below 2*nextd+2 per axis) for parity tests the CPU oracle can finish quickly."""
    def sc(v, lo):
        return max(lo, int(round(v * scale)))
    if n == 1:   # shipped example: 58x58x43 8-node, f=0.1 Hz; GPML Zhou as shipped, or Dirichlet (Q14)
        nextd = 4 if scale == 1.0 else 2
        d = 0 if dirichlet is None else dirichlet
        return build_model("config1_shipped_linear", sc(58, 2 * nextd + 2), sc(58, 2 * nextd + 2), 8, 1990., 1990., 2000.,
                           nextd, sc(25, 2), sc(10, 1), dirichlet=d, gpml_sch=1, a0=1., b0=1., nn=2., freqs=(0.1,))
    if n == 2:   # 40x40x30 20-node, layered + block, GPML Fang, f=1 Hz
        nextd = 4 if scale == 1.0 else 2
        return build_model("config2_quadratic_gpml_fang", sc(40, 2 * nextd + 2), sc(40, 2 * nextd + 2), 20, 1000., 1000., 1000.,
                           nextd, sc(16, 2), sc(6, 1), dirichlet=0 if dirichlet is None else dirichlet, gpml_sch=0,
                           a0=1., b0=1., nn=2., freqs=(1.0,), sigma_fn=_layered((4000., 4000., 1000., 5000.)))
    if n == 3:   # 24x24x18 27-node, full anisotropic sigma, GPML Zhou, f=1 Hz
        nextd = 4 if scale == 1.0 else 2
        return build_model("config3_lagrange_aniso_gpml_zhou", sc(24, 2 * nextd + 2), sc(24, 2 * nextd + 2), 27, 1000., 1000., 1000.,
                           nextd, sc(7, 2), sc(3, 1), dirichlet=0 if dirichlet is None else dirichlet, gpml_sch=1,
                           a0=1., b0=1., nn=2., freqs=(1.0,), sigma_fn=_layered(), aniso=True, seed=20141)
    if n == 4:   # 100x100x60 linear, 32 frequencies 1e-3..1e3 Hz
        nextd = 4 if scale == 1.0 else 2
        return build_model("config4_sweep_linear", sc(100, 2 * nextd + 2), sc(100, 2 * nextd + 2), 8, 1000., 1000., 1000.,
                           nextd, sc(36, 2), sc(16, 1), dirichlet=0 if dirichlet is None else dirichlet, gpml_sch=1,
                           a0=1., b0=1., nn=2., freqs=np.logspace(-3, 3, 32), sigma_fn=_layered())
    if n == 5:   # 400x400x200 linear with topography, GPML Fang, f=1 Hz
        nextd = 4 if scale == 1.0 else 2
        return build_model("config5_large_topography", sc(400, 2 * nextd + 2), sc(400, 2 * nextd + 2), 8, 250., 250., 250.,
                           nextd, sc(140, 2), sc(52, 1), dirichlet=0 if dirichlet is None else dirichlet, gpml_sch=0,
                           a0=1., b0=1., nn=2., freqs=(1.0,), sigma_fn=_layered(), topo_amp=300.0 if scale == 1.0 else 60.0)
    raise ValueError(n)


CONFIG_NAMES = {1: "config1_shipped_linear", 2: "config2_quadratic_gpml_fang", 3: "config3_lagrange_aniso_gpml_zhou",
                4: "config4_sweep_linear", 5: "config5_large_topography"}


def config_name(n: int) -> str:
    """Name of BASELINE.json configs[n-1] without building the mesh."""
    return CONFIG_NAMES[n]


def config5_submesh() -> Model:
    """The parity mesh of BASELINE configs[4]: 40x40x20 linear elements with the FULL topography amplitude
    (300 m on every interface, tapering to 0 at the bottom / top planes), nextd = 4, GPML Fang, f = 1 Hz."""
    return build_model("config5_submesh_40x40x20", 40, 40, 8, 250., 250., 250., 4, 9, 3, dirichlet=0, gpml_sch=0,
                       a0=1., b0=1., nn=2., freqs=(1.0,), sigma_fn=_layered(), topo_amp=300.0)


def brick_single_element(mn: int, hx=2000.0, hy=1500.0, hz=1000.0, top_shift=None) -> Model:
    """One element whose nodes are ``nf_nr(l,:)*(hx,hy,hz)/2`` (SURVEY App. B item 4 pins).
    ``top_shift`` raises the four top corner nodes 5..8 of an 8-node element (Q5 pin)."""
    me, nord = ELEMENT_TYPES[mn]
    ln = np.array([-0.5, 0.5]) if nord == 2 else np.array([-0.5, 0.0, 0.5])
    g_xp, g_yp, zc = ln * hx, ln * hy, ln * hz
    n = ln.size
    Z = np.broadcast_to(zc, (n, n, n)).copy()
    if top_shift is not None:
        # local nodes 5..8 sit at (xi,eta) = (+,-),(+,+),(-,+),(-,-)  (n_fem.f90:114-116)
        Z[1, 0, 1] += top_shift[0]; Z[1, 1, 1] += top_shift[1]; Z[0, 1, 1] += top_shift[2]; Z[0, 0, 1] += top_shift[3]
    npt = n ** 3
    sig = np.zeros((npt, 6)); sig[:, [0, 3, 5]] = 0.01
    mask = np.zeros((npt, 6), bool)                               # purely real sigma for the pins
    g_mu = np.zeros((npt, 6)); g_mu[:, [0, 3, 5]] = MU0
    return Model(name=f"brick_mn{mn}", g_nx=2, g_ny=2, g_nz=2, mn=mn, nextd=1, nzl_top=1, dirichlet=0, gpml_sch=1,
                 a0=1., b0=1., nn=2., g_xp=g_xp.copy(), g_yp=g_yp.copy(), g_zp=np.ascontiguousarray(Z.reshape(-1)),
                 g_mu=g_mu, sigma_re=sig, sigma_im_mask=mask, freqs=np.array([0.1]))

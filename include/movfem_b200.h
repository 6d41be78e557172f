/*
 * movfem_b200.h -- C ABI of the B200-native vector-FEM assembly for MoVFEM_3DMT.
 *
 * This is the drop-in boundary.  It replaces, in the reference,
 *
 *   MoVFEM_3DMT.f90:80-97   call global_vfem(irn,jcn,a,rhs) + find_zeros/rem_zeros
 *   global_assembly.f90:38-39  ga_cgne / ga_nzindx  (DOF numbering + pattern, once per mesh)
 *
 * and nothing else: the Fortran driver, geometry, readers and the ZMUMPS solve stay.
 * All entry points are extern "C", take plain pointers and sizes, return 0 on success or
 * a negative MOVFEM_E_* code (the reference's `stop`s at global_assembly.f90:56-57,109-111,
 * n_fem.f90:374-377, problem.f90:260-271 become return codes).  Arrays use the Fortran
 * caller's conventions: column-major, 1-based index VALUES, caller-owned host memory.
 *
 * The Fortran-side binding (ISO_C_BINDING) is in movfem_b200/fortran/movfem_cuda.f90 and
 * described in INTEGRATION.md.
 *
 * Threading: a handle is used by one host thread at a time.  Handles of DIFFERENT element types (8 / 20 / 27 nodes) on the
 * same device share one __constant__ operand table: assemble them from one host thread (the switch drains the device first),
 * not concurrently from several.  Handles of the same element type, and handles on different devices, are independent.
 */
#ifndef MOVFEM_B200_H
#define MOVFEM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* error codes */
#define MOVFEM_OK                 0
#define MOVFEM_E_BADARG         (-1)  /* inconsistent descriptor / null pointer            */
#define MOVFEM_E_CUDA           (-2)  /* CUDA runtime error (see movfem_last_error)        */
#define MOVFEM_E_SINGULAR_JAC   (-3)  /* n_fem.f90:374-377  'no transformation!! nf_det=0' */
#define MOVFEM_E_SINGULAR_MODEL (-4)  /* problem.f90:260-271 'no sigma/mu inversion'       */
#define MOVFEM_E_NOGPU          (-5)  /* no CUDA device: there is NO CPU fallback          */
#define MOVFEM_E_CAPACITY       (-6)  /* caller array too small                            */
#define MOVFEM_E_UNSUPPORTED    (-7)  /* configuration outside what the driver hard-codes  */

/* assembly output modes */
#define MOVFEM_MODE_T2  0  /* what ZMUMPS receives: upper triangle, row-major sorted, values
                              rounded through float32 (global_assembly.f90:157,166,169),
                              exact zeros stripped (global_assembly.f90:123-150)           */
#define MOVFEM_MODE_T1  1  /* tap before ga_sort_sparse: same ordering and pattern
                              (structural upper triangle, nothing stripped), double values */
#define MOVFEM_MODE_KEEP_PATTERN 0x100  /* OR-ed into the mode of movfem_assemble: the caller's irn/jcn still hold what
                              the previous movfem_assemble on this handle delivered (the structural pattern is static
                              across frequencies), so they are re-sent only if the zero strip changed the delivered set.
                              Saves 8 of the 24 B/entry on the host link.  The reference reallocates irn/jcn every
                              frequency (MoVFEM_3DMT.f90:78,119); hoist that allocation to use this.             */

/*
 * Mesh / problem descriptor.  Every field is the reference module variable of the same
 * name (geometry.f90:17-26, boundary_conds.f90:15-25, problem.f90:18, v_fem.f90:16,
 * n_fem.f90:14) so the Fortran shim fills it by plain assignment.
 */
typedef struct movfem_desc {
    int32_t g_nx, g_ny, g_nz;   /* grid LINES per axis: elements are (g_nx-1)(g_ny-1)(g_nz-1) */
    int32_t nord;               /* g_nordx=g_nordy=g_nordz: 2 (8-node) or 3 (20/27-node)      */
    int32_t mn;                 /* nf_mn: 8, 20 or 27 nodes per element                       */
    int32_t me;                 /* vf_me: 12, 36 or 54 edge DOFs per element                  */
    int32_t nextd;              /* geometry.f90 nextd: extension / GPML layers per side       */
    int32_t nzl_top;            /* g_nzl(g_nsf): element layers of the top extension          */
    int32_t dirichlet;          /* boundary_conds.f90:15  1 = Dirichlet, 0 = GPML             */
    int32_t bd_inimod;          /* Dirichlet boundary model (PARAM.INP line 7): 1 zero, 2 homogeneous, 3 layered */
    int32_t gpml_sch;           /* 0 = Fang 1996, 1 = Zhou 2012 (boundary_conds.f90:100-108)  */
    int32_t sym;                /* global_assembly.f90:18; the driver hard-codes 1            */
    int32_t ndir;               /* problem.f90 ndir; the driver hard-codes 2                  */
    int32_t pe_sch;             /* problem.f90 pe_sch; the driver hard-codes 1 (secondary E)  */
    double  a0, b0, nn;         /* GPML constants (PARAM.INP last line)                       */
    const double *g_xp;         /* (g_nnx)   x of node lines                                  */
    const double *g_yp;         /* (g_nny)   y of node lines                                  */
    const double *g_zp;         /* (g_npt)   z of every node, id=(ii-1)*g_nyz+(jj-1)*g_nnz+kk */
    const double *g_mu;         /* (6,g_npt) real permeability tensor 11,12,13,22,23,33       */
    /* element slab owned by this handle (multi-GPU slab sharding); 1-based,
       inclusive, in the reference's ie index.  0,0 = whole mesh.                            */
    int32_t ie_lo, ie_hi;
    /* Dirichlet boundary models 2 / 3 (boundary_conds.f90:188-250,392-598, arguments of bd_setmodel,
       MoVFEM_3DMT.f90:388): the primary field of a homogeneous / layered earth on the side faces          */
    double  g_ztop;             /* geometry.f90:394  lowest point of the topography interface              */
    double  bd_hsigma;          /* PARAM.INP: conductivity of the half-space (bd_inimod = 2)               */
    int32_t bd_nl;              /* PARAM.INP: number of layers (bd_inimod = 3), 1..16                       */
    int32_t bd_pad;
    double  bd_lsigma[16];      /* conductivity of layers 1..nl                                            */
    double  bd_ldz[16];         /* thickness of layers 1..nl-1, as handed to bd_setmodel                    */
} movfem_desc;

typedef struct movfem_handle movfem_handle;

/* phase timings of the last movfem_assemble call, milliseconds (CUDA events) */
typedef struct movfem_stats {
    double ms_h2d, ms_node, ms_element, ms_gather, ms_finalize, ms_d2h, ms_total;
    int64_t nz;           /* entries delivered                                  */
    int64_t launches;     /* kernels launched by the call                       */
    double ms_geometry;   /* part of ms_element: geometry_kernel launches        */
    double ms_contract;   /* part of ms_element: contract_kernel launches        */
    double ms_exact;      /* part of ms_element: exact_kernel launches (reference-order re-evaluation of residue pairs) */
    int64_t nflagged;     /* (element, pair)s re-evaluated in the reference's operation order                          */
    double ms_fused;      /* part of ms_element: fused12_kernel (linear elements: geometry + contraction + RHS in one)  */
} movfem_stats;

/* Create: uploads the mesh once, builds gne + pattern ON THE DEVICE
   (replaces ga_init -> ga_cgne/ga_nzindx, global_assembly.f90:26-41).                        */
int movfem_create(const movfem_desc *desc, int device, movfem_handle **out);
void movfem_destroy(movfem_handle *h);

/* nne, nnze (full structural pattern, what MoVFEM_3DMT.f90:72-78 allocates) and the
   structural upper-triangle count (capacity the graft actually needs).                      */
int movfem_sizes(const movfem_handle *h, int32_t *nne, int64_t *nnze_full, int64_t *nz_upper);

/* Rows owned by the handle: 1-based first row and count.  The whole matrix unless the descriptor asked for an
   x-slab (ie_lo..ie_hi): a slab handle owns the rows whose first-encounter element lies in the slab -- a contiguous
   range, because DOFs are numbered in (ie,je,ke) order (global_assembly.f90:237-296) -- computes the +x neighbour
   layer itself (no exchange) and delivers only those rows; movfem_sizes then reports the LOCAL nz_upper and 0 for
   nnze_full, and movfem_assemble fills only rows row_lo..row_lo+nrows-1 of each RHS column.               */
int movfem_slab_rows(const movfem_handle *h, int32_t *row_lo, int32_t *nrows);

/* gne(ne,me), column-major, Fortran values (1-based, -face for Dirichlet edges):
   global_assembly.f90:183-195.  solution.f90:331-336 consumes it after the solve.           */
int movfem_get_gne(const movfem_handle *h, int32_t *gne);

/* Structural upper-triangle pattern in delivery order (row-major, 1-based). */
int movfem_get_pattern(const movfem_handle *h, int32_t *irn, int32_t *jcn);

/*
 * One frequency: replaces MoVFEM_3DMT.f90:82-97.
 *   freq_index  1-based position in the reference's sequential frequency loop (element
 *               (1,1,1) sees stale GPML flags from the previous frequency)
 *   omega       geometry.f90 omega = 2*pi*f
 *   g_sigma     (6,g_npt) complex128, re-read every call
 *   irn,jcn,a   capacity >= nz_upper entries;  rhs  ndir*nne entries
 *   nz_out      number of triplets delivered (mumps_par%nz)
 */
int movfem_assemble(movfem_handle *h, int32_t freq_index, double omega,
                    const double *g_sigma /* complex128 as re,im pairs */,
                    int32_t *irn, int32_t *jcn, double *a, double *rhs,
                    int64_t *nz_out, int32_t mode);

/*
 * Device-resident variant used for kernel-only timing and for device consumers
 * (SURVEY 8f-3): g_sigma_dev is a device pointer; results stay on the device and are
 * reachable through movfem_device_result.  No host copies, asynchronous on the stream.
 */
int movfem_assemble_device(movfem_handle *h, int32_t freq_index, double omega,
                           const double *g_sigma_dev, int32_t mode);
int movfem_device_result(const movfem_handle *h, const int32_t **irn, const int32_t **jcn,
                         const double **a, const double **rhs, int64_t *nz /* syncs */);
/* CSR view of the last device result for a GPU sparse direct solver: the delivered triplets are the
   upper triangle sorted by (row, col), i.e. CSR minus the row pointers; rowptr[0..nrows] (device, 0-based offsets
   into irn/jcn/a, rows local to the handle) completes it: row r holds entries rowptr[r] .. rowptr[r+1]-1, its
   column indices are jcn (1-based) and its values a.  Valid until the next assemble on the handle.           */
int movfem_device_csr(const movfem_handle *h, const int64_t **rowptr, int32_t *nrows);
/* A device consumer of that CSR view: y = A x (complex128 device vectors of nne entries) for the complex symmetric matrix whose
   upper triangle is the last device result -- the residual / refinement step of a GPU sparse solver.  ms_device (optional):
   device time of the two kernels.  Whole-mesh handles only (a slab handle returns MOVFEM_E_UNSUPPORTED).               */
int movfem_device_spmv(const movfem_handle *h, const double *x_dev, double *y_dev, double *ms_device);

/* Forget the cached K_e/M_e of the unstretched elements: the next assemble recomputes every
   element (what a single-frequency run does).  A sweep keeps them.            */
int movfem_reset_cache(movfem_handle *h);

/* Measured FP64 FMA-loop throughput of the device in TFLOP/s (roofline denominator). */
int movfem_fp64_peak(int device, double *tflops);

/*
 * Geomodel -> grid nodes: replaces geometry.f90:801-970 innermodel_gqg with min_dd_inner (:975-1031,
 * a serial O(npt x model cells) nearest-neighbour search) and assign_model (:1037-1085).  Arguments are those of the
 * reference call at geometry.f90:99 (after coord_transform) plus the mesh it reads from module geometry.
 */
typedef struct movfem_geomodel {
    int32_t mx, my, mz;          /* read_input.f90 in_mx, in_my, in_mz: model grid                               */
    int32_t isigma, imu;         /* number of conductivity / permeability tensor components given (1..9)         */
    int32_t nzl_air;             /* g_nzl(g_nsf-1): element layers of the air below the top extension            */
    int32_t ijsigma[9][2];       /* ijsigma(i,1:2): (row, col) of component i                                    */
    int32_t ijmu[9][2];
    const double *xm, *ym;       /* (mx), (my)                                                                   */
    const double *zm;            /* (mx*my*mz), idd=(im-1)*my*mz+(jm-1)*mz+km                                    */
    const double *sigma;         /* (isigma, mx*my*mz) column-major                                              */
    const double *mu;            /* (imu, mx*my*mz) column-major, relative permeability                          */
} movfem_geomodel;
/* mesh: g_nx,g_ny,g_nz,nord,nextd,nzl_top,g_xp,g_yp,g_zp of the descriptor are read.  omega = 2*pi*g_freq(1)
   (geometry.f90:73).  g_sigma (6,g_npt) complex128 and g_mu (6,g_npt) are written as innermodel_gqg leaves them.
   ms_device (optional): device time of the kernels in milliseconds.                                             */
int movfem_geo_innermodel(const movfem_desc *mesh, int32_t device, const movfem_geomodel *gm, double omega,
                          double *g_sigma, double *g_mu, double *ms_device);

/* stream the handle launches on (cudaStream_t as void*); set before assembling. */
int movfem_set_stream(movfem_handle *h, void *cuda_stream);
int movfem_get_stats(const movfem_handle *h, movfem_stats *out);
const char *movfem_last_error(const movfem_handle *h);
const char *movfem_version(void);

/* debug / parity tap: element matrices of one element as the kernels compute them
   (K_e, M_e lower-by-local-index packed me*(me+1)/2, b_e me*ndir complex).                  */
int movfem_debug_element(movfem_handle *h, int32_t ide /*1-based*/, double *Ke, double *Me,
                         double *be);

#ifdef __cplusplus
}
#endif
#endif /* MOVFEM_B200_H */

/*
 * abi_harness.c -- plain-C caller of include/movfem_b200.h (test infrastructure).
 *
 * Compiled as C99 with -Wall -Wextra -Werror -pedantic by tests/test_c_harness.py: proves that the header is a C
 * header (not only C++), that the library links from C with nothing but -lmovfem_b200, and reports the struct layouts
 * the C compiler derives so the ctypes mirror (movfem_b200/abi.py) and the Fortran bind(C) types
 * (movfem_b200/fortran/movfem_cuda.f90) can be compared with them.
 *
 * It drives the boundary the way MoVFEM_3DMT.f90 would: create -> sizes -> get_gne -> assemble ->
 * destroy on a 2x2x2 brick of 8-node elements with Fortran conventions (column-major, 1-based values, caller-owned
 * arrays).  Without a CUDA device movfem_create must return MOVFEM_E_NOGPU (there is no CPU fallback) and the
 * harness prints that; with a device it assembles and prints nne / nz.
 */
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "movfem_b200.h"

#define OFF(T, f) printf("offset " #T "." #f " %lu\n", (unsigned long)offsetof(T, f))

int main(int argc, char **argv)
{
    enum { NL = 3, NE = 8, ME = 12, NPT = NL * NL * NL };   /* 3 grid lines per axis, linear elements */
    double xp[NL] = {0.0, 1000.0, 2100.0}, yp[NL] = {0.0, 900.0, 2000.0};
    double zp[NPT], mu[6 * NPT], sigma[2 * 6 * NPT];
    movfem_desc d;
    movfem_handle *h = NULL;
    int rc, i, j, k;

    printf("sizeof movfem_desc %lu\n", (unsigned long)sizeof(movfem_desc));
    printf("sizeof movfem_stats %lu\n", (unsigned long)sizeof(movfem_stats));
    printf("sizeof movfem_geomodel %lu\n", (unsigned long)sizeof(movfem_geomodel));
    OFF(movfem_desc, g_nx); OFF(movfem_desc, nzl_top); OFF(movfem_desc, dirichlet); OFF(movfem_desc, pe_sch);
    OFF(movfem_desc, a0); OFF(movfem_desc, nn); OFF(movfem_desc, g_xp); OFF(movfem_desc, g_mu);
    OFF(movfem_desc, ie_lo); OFF(movfem_desc, g_ztop); OFF(movfem_desc, bd_hsigma); OFF(movfem_desc, bd_nl);
    OFF(movfem_desc, bd_lsigma); OFF(movfem_desc, bd_ldz);
    OFF(movfem_stats, nz); OFF(movfem_stats, launches); OFF(movfem_stats, ms_contract);
    OFF(movfem_geomodel, nzl_air); OFF(movfem_geomodel, ijsigma); OFF(movfem_geomodel, ijmu);
    OFF(movfem_geomodel, xm); OFF(movfem_geomodel, mu);
    printf("version %s\n", movfem_version());
    if (argc > 1 && strcmp(argv[1], "layout") == 0) return 0;   /* layouts only: no device call */

    /* node id = (ii-1)*g_nyz + (jj-1)*g_nnz + kk, z fastest (n_fem.f90:66-102) */
    for (i = 0; i < NL; ++i)
        for (j = 0; j < NL; ++j)
            for (k = 0; k < NL; ++k) {
                const int id = (i * NL + j) * NL + k;
                zp[id] = 800.0 * k + 15.0 * i - 10.0 * j;
                memset(&mu[6 * id], 0, 6 * sizeof(double));
                mu[6 * id + 0] = mu[6 * id + 3] = mu[6 * id + 5] = 1.25663706143591729e-6;
                memset(&sigma[12 * id], 0, 12 * sizeof(double));
                sigma[12 * id + 0] = sigma[12 * id + 6] = sigma[12 * id + 10] = 0.01;   /* Re of xx, yy, zz */
            }

    memset(&d, 0, sizeof d);
    d.g_nx = d.g_ny = d.g_nz = NL;
    d.nord = 2; d.mn = 8; d.me = ME; d.nextd = 1; d.nzl_top = 1;
    d.dirichlet = 1; d.bd_inimod = 1; d.gpml_sch = 0; d.sym = 1; d.ndir = 2; d.pe_sch = 1;
    d.a0 = 1.0; d.b0 = 1.0; d.nn = 2.0;
    d.g_xp = xp; d.g_yp = yp; d.g_zp = zp; d.g_mu = mu;
    d.bd_nl = 1;

    rc = movfem_create(&d, 0, &h);
    printf("create %d\n", rc);
    if (rc == MOVFEM_E_NOGPU) {
        if (h != NULL) { printf("handle must stay NULL on failure\n"); return 1; }
        return 0;
    }
    if (rc != MOVFEM_OK) return 1;
    {
        int32_t nne = 0, *gne, *irn, *jcn;
        int64_t nnze = 0, nzu = 0, nz = 0;
        double *a, *rhs;
        rc = movfem_sizes(h, &nne, &nnze, &nzu);
        printf("sizes %d nne %d nnze %ld nz_upper %ld\n", rc, (int)nne, (long)nnze, (long)nzu);
        gne = (int32_t *)malloc(sizeof(int32_t) * NE * ME);
        irn = (int32_t *)malloc(sizeof(int32_t) * (size_t)nzu);
        jcn = (int32_t *)malloc(sizeof(int32_t) * (size_t)nzu);
        a = (double *)malloc(sizeof(double) * 2 * (size_t)nzu);
        rhs = (double *)malloc(sizeof(double) * 2 * 2 * (size_t)nne);
        rc = movfem_get_gne(h, gne);
        printf("get_gne %d gne(1,1) %d\n", rc, (int)gne[0]);
        rc = movfem_assemble(h, 1, 6.283185307179586 * 0.1, sigma, irn, jcn, a, rhs, &nz, MOVFEM_MODE_T2);
        printf("assemble %d nz %ld\n", rc, (long)nz);
        /* 2x2x2 bricks, all faces Dirichlet: the six edges at the centre node are the unknowns; 18 upper-triangle entries
           (the CPU oracle's figures for this mesh) */
        if (rc == MOVFEM_OK && (nne != 6 || nzu != 18 || nz != 18)) { printf("unexpected sizes\n"); rc = 1; }
        if (rc != MOVFEM_OK) printf("error %s\n", movfem_last_error(h));
        for (i = 0; i + 1 < (int)nz; ++i)   /* delivered order: row-major, upper triangle, 1-based */
            if (irn[i] > irn[i + 1] || (irn[i] == irn[i + 1] && jcn[i] >= jcn[i + 1]) || jcn[i] < irn[i] || irn[i] < 1) {
                printf("order violated at %d\n", i);
                rc = 1;
            }
        free(gne); free(irn); free(jcn); free(a); free(rhs);
        movfem_destroy(h);
    }
    if (rc == MOVFEM_OK) printf("OK\n");
    return rc == MOVFEM_OK ? 0 : 1;
}

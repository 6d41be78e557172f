/*
 * movfem_b200/csrc/ref_element.h -- reference-element tables of the PRODUCT (host side).
 *
 * Nodal (8 / 20 / 27) and mixed-order edge (12 / 36 / 54) shape functions of MoVFEM_3DMT,
 * evaluated once per handle at the Gauss points and uploaded to the device as constant
 * tables (the Gauss points are fixed, so N, dN/dxi, phi, dphi/dxi are element independent).
 * The expressions keep the evaluation order of n_fem.f90:138-317 and v_fem.f90:78-466 and this
 * file is compiled with -ffp-contract=off, so the tables carry the same bits the Fortran
 * functions return -- including the 8-node dN/dzeta typo of n_fem.f90:193 and the
 * default-real Gauss literals of integration.f90:297-299,400-402.
 * tests/test_tables.py compares them bit for bit with the independent copy in oracle/.
 */
#ifndef MOVFEM_REF_ELEMENT_H
#define MOVFEM_REF_ELEMENT_H

namespace movfem {

// ---- n_fem.f90:36-59 (nord = g_nordx = g_nordy = g_nordz) ---------------------------------
inline void node_offsets(int mn, int nord, int *i1, int *j1, int *k1) {
    const int g = nord;
    if (mn == 8) {
        const int a[8] = {g, g, 1, 1, g, g, 1, 1};
        const int b[8] = {1, g, g, 1, 1, g, g, 1};   // first 8 of the 20-entry constructor, n_fem.f90:39-40
        const int c[8] = {1, 1, 1, 1, g, g, g, g};
        for (int n = 0; n < 8; ++n) { i1[n] = a[n]; j1[n] = b[n]; k1[n] = c[n]; }
    } else if (mn == 20) {
        const int a[20] = {g, g, 1, 1, g, g, 1, 1, g, 2, 1, 2, g, 2, 1, 2, g, g, 1, 1};
        const int b[20] = {1, g, g, 1, 1, g, g, 1, 2, g, 2, 1, 2, g, 2, 1, 1, g, g, 1};
        const int c[20] = {1, 1, 1, 1, g, g, g, g, 1, 1, 1, 1, g, g, g, g, 2, 2, 2, 2};
        for (int n = 0; n < 20; ++n) { i1[n] = a[n]; j1[n] = b[n]; k1[n] = c[n]; }
    } else {
        const int a[27] = {g, g, 1, 1, g, g, 1, 1, g, 2, 1, 2, g, 2, 1, 2, g, g, 1, 1, g, 2, 1, 2, 2, 2, 2};
        const int b[27] = {1, g, g, 1, 1, g, g, 1, 2, g, 2, 1, 2, g, 2, 1, 1, g, g, 1, 2, g, 2, 1, 2, 2, 2};
        const int c[27] = {1, 1, 1, 1, g, g, g, g, 1, 1, 1, 1, g, g, g, g, 2, 2, 2, 2, 2, 2, 2, 2, 1, g, 2};
        for (int n = 0; n < 27; ++n) { i1[n] = a[n]; j1[n] = b[n]; k1[n] = c[n]; }
    }
}

// ---- n_fem.f90:113-134 ------------------------------------------------------------------
inline void node_ref_coords(int mn, double (*nr)[3]) {
    static const int x27[27] = {1, 1, -1, -1, 1, 1, -1, -1, 1, 0, -1, 0, 1, 0, -1, 0, 1, 1, -1, -1, 1, 0, -1, 0, 0, 0, 0};
    static const int y27[27] = {-1, 1, 1, -1, -1, 1, 1, -1, 0, 1, 0, -1, 0, 1, 0, -1, -1, 1, 1, -1, 0, 1, 0, -1, 0, 0, 0};
    static const int z27[27] = {-1, -1, -1, -1, 1, 1, 1, 1, -1, -1, -1, -1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0, -1, 1, 0};
    // the 8- and 20-node tables are prefixes of the 27-node one (n_fem.f90:114-123)
    for (int n = 0; n < mn; ++n) { nr[n][0] = x27[n]; nr[n][1] = y27[n]; nr[n][2] = z27[n]; }
}

// ---- v_fem.f90:491-504 (1-based node, direction) -----------------------------------------
inline void edge_dir_table(int me, int *node, int *dir) {
    if (me == 12) {
        const int a[12] = {4, 4, 8, 3, 4, 8, 3, 7, 1, 1, 5, 2};
        const int d[12] = {3, 2, 2, 3, 1, 1, 1, 1, 3, 2, 2, 3};
        for (int e = 0; e < 12; ++e) { node[e] = a[e]; dir[e] = d[e]; }
    } else if (me == 36) {
        const int a[36] = {4, 8, 4, 3, 8, 7, 3, 7, 4, 1, 8, 5, 3, 2, 7, 6, 1, 5, 1, 2, 5, 6, 2,
                           6, 12, 20, 12, 16, 10, 17, 11, 20, 11, 15, 19, 9};
        const int d[36] = {3, 3, 2, 2, 2, 2, 3, 3, 1, 1, 1, 1, 1, 1, 1, 1, 3, 3, 2, 2, 2, 2, 3,
                           3, 3, 2, 2, 2, 3, 2, 3, 1, 1, 1, 1, 3};
        for (int e = 0; e < 36; ++e) { node[e] = a[e]; dir[e] = d[e]; }
    } else {
        const int a[54] = {4, 8, 4, 20, 8, 11, 15, 3, 19, 7, 3, 7, 4, 20, 8, 11, 15, 3, 19,
                           7, 12, 16, 12, 10, 16, 14, 10, 14, 1, 17, 5, 9, 13, 2, 18, 6, 1, 5, 1, 17, 5, 9, 13, 2, 18,
                           6, 2, 6, 25, 26, 24, 22, 23, 21};
        const int d[54] = {3, 3, 2, 2, 2, 3, 3, 2, 2, 2, 3, 3, 1, 1, 1, 1, 1, 1, 1, 1, 3, 3, 2,
                           2, 2, 2, 3, 3, 1, 1, 1, 1, 1, 1, 1, 1, 3, 3, 2, 2, 2, 3, 3, 2, 2, 2, 3, 3, 3, 3, 2, 2, 1, 1};
        for (int e = 0; e < 54; ++e) { node[e] = a[e]; dir[e] = d[e]; }
    }
}

// node class used by the select-case ladders (1-based node id)
enum NodeClass { NC_CORNER, NC_MIDX /*10,12,14,16*/, NC_MIDY /*9,11,13,15*/, NC_MIDZ /*17..20*/,
                 NC_FACEX /*21,23*/, NC_FACEY /*22,24*/, NC_FACEZ /*25,26*/, NC_CENTRE /*27*/ };
inline NodeClass node_class(int i) {
    if (i <= 8) return NC_CORNER;
    if (i <= 16) return (i % 2 == 0) ? NC_MIDX : NC_MIDY;
    if (i <= 20) return NC_MIDZ;
    if (i == 21 || i == 23) return NC_FACEX;
    if (i == 22 || i == 24) return NC_FACEY;
    if (i <= 26) return NC_FACEZ;
    return NC_CENTRE;
}

struct Shape {
    int mn;
    double nr[27][3];
    explicit Shape(int mn_) : mn(mn_) { node_ref_coords(mn, nr); }

    // n_fem.f90:138-178; i is 1-based
    double nf_ln(int i, double xi, double eta, double zeta) const {
        const double a = nr[i - 1][0], b = nr[i - 1][1], c = nr[i - 1][2];
        if (mn == 8) return ((1 + a * xi) * (1 + b * eta) * (1 + c * zeta)) / 8.0;
        if (mn == 20) {
            switch (node_class(i)) {
            case NC_CORNER: return ((1 + a * xi) * (1 + b * eta) * (1 + c * zeta) * (a * xi + b * eta + c * zeta - 2)) / 8.0;
            case NC_MIDX:   return ((1 - xi * xi) * (1 + b * eta) * (1 + c * zeta)) / 4.0;
            case NC_MIDY:   return ((1 + a * xi) * (1 - eta * eta) * (1 + c * zeta)) / 4.0;
            case NC_MIDZ:   return ((1 + a * xi) * (1 + b * eta) * (1 - zeta * zeta)) / 4.0;
            default: return 0.0;
            }
        }
        switch (node_class(i)) {
        case NC_CORNER: return ((1 + a * xi) * (1 + b * eta) * (1 + c * zeta) * (a * xi * b * eta * c * zeta)) / 8.0;
        case NC_MIDX:   return ((1 - xi * xi) * (1 + b * eta) * (1 + c * zeta) * b * eta * c * zeta) / 4.0;
        case NC_MIDY:   return ((1 + a * xi) * (1 - eta * eta) * (1 + c * zeta) * a * xi * c * zeta) / 4.0;
        case NC_MIDZ:   return ((1 + a * xi) * (1 + b * eta) * (1 - zeta * zeta) * a * xi * b * eta) / 4.0;
        case NC_FACEX:  return (1 - eta * eta) * (1 - zeta * zeta) * (1 + a * xi) * a * xi / 2.0;
        case NC_FACEY:  return (1 - xi * xi) * (1 - zeta * zeta) * (1 + b * eta) * b * eta / 2.0;
        case NC_FACEZ:  return (1 - xi * xi) * (1 - eta * eta) * (1 + c * zeta) * c * zeta / 2.0;
        default:        return (1 - xi * xi) * (1 - eta * eta) * (1 - zeta * zeta);
        }
    }

    // n_fem.f90:181-317; d = 1,2,3; i 1-based
    double nf_dln_dxi(int d, int i, double xi, double eta, double zeta) const {
        const double a = nr[i - 1][0], b = nr[i - 1][1], c = nr[i - 1][2];
        if (mn == 8) {
            if (d == 1) return (a * (1 + b * eta) * (1 + c * zeta)) / 8.0;
            if (d == 2) return ((1 + a * xi) * b * (1 + c * zeta)) / 8.0;
            return ((1 + a) * (1 + b * eta) * c) / 8.0;            // n_fem.f90:193 -- Q5: no *xi
        }
        if (mn == 20) {
            switch (node_class(i)) {
            case NC_CORNER:
                if (d == 1) return ((1 + b * eta) * (1 + c * zeta) * (a * (2 * a * xi + b * eta + c * zeta - 1))) / 8.0;
                if (d == 2) return ((1 + a * xi) * (1 + c * zeta) * (b * (2 * b * eta + a * xi + c * zeta - 1))) / 8.0;
                return ((1 + b * eta) * (1 + a * xi) * (c * (2 * c * zeta + b * eta + a * xi - 1))) / 8.0;
            case NC_MIDX:
                if (d == 1) return -xi * (1 + b * eta) * (1 + c * zeta) / 2.0;
                if (d == 2) return b * (1 - xi * xi) * (1 + c * zeta) / 4.0;
                return c * (1 - xi * xi) * (1 + b * eta) / 4.0;
            case NC_MIDY:
                if (d == 1) return a * (1 - eta * eta) * (1 + c * zeta) / 4.0;
                if (d == 2) return -eta * (1 + a * xi) * (1 + c * zeta) / 2.0;
                return c * (1 - eta * eta) * (1 + a * xi) / 4.0;
            case NC_MIDZ:
                if (d == 1) return a * (1 - zeta * zeta) * (1 + b * eta) / 4.0;
                if (d == 2) return b * (1 - zeta * zeta) * (1 + a * xi) / 4.0;
                return -zeta * (1 + a * xi) * (1 + b * eta) / 2.0;
            default: return 0.0;
            }
        }
        switch (node_class(i)) {
        case NC_CORNER:
            if (d == 1) return (a * (1 + 2 * a * xi) * (1 + b * eta) * (1 + c * zeta) * (b * eta * c * zeta)) / 8.0;
            if (d == 2) return (b * (1 + a * xi) * (1 + 2 * b * eta) * (1 + c * zeta) * (a * xi * c * zeta)) / 8.0;
            return (c * (1 + a * xi) * (1 + b * eta) * (1 + 2 * c * zeta) * (a * xi * b * eta)) / 8.0;
        case NC_MIDX:
            if (d == 1) return (-xi * (1 + b * eta) * (1 + c * zeta) * b * eta * c * zeta) / 2.0;
            if (d == 2) return (b * (1 - xi * xi) * (1 + 2 * b * eta) * (1 + c * zeta) * c * zeta) / 4.0;
            return (c * (1 - xi * xi) * (1 + b * eta) * (1 + 2 * c * zeta) * b * eta) / 4.0;
        case NC_MIDY:
            if (d == 1) return (a * (1 + 2 * a * xi) * (1 - eta * eta) * (1 + c * zeta) * c * zeta) / 4.0;
            if (d == 2) return (-eta * (1 + a * xi) * (1 + c * zeta) * a * xi * c * zeta) / 2.0;
            return (c * (1 + a * xi) * (1 - eta * eta) * (1 + 2 * c * zeta) * a * xi) / 4.0;
        case NC_MIDZ:
            if (d == 1) return (a * (1 + 2 * a * xi) * (1 + b * eta) * (1 - zeta * zeta) * b * eta) / 4.0;
            if (d == 2) return (b * (1 + a * xi) * (1 + 2 * b * eta) * (1 - zeta * zeta) * a * xi) / 4.0;
            return (-zeta * (1 + a * xi) * (1 + b * eta) * a * xi * b * eta) / 2.0;
        case NC_FACEX:
            if (d == 1) return a * (1 - eta * eta) * (1 - zeta * zeta) * (1 + 2 * a * xi) / 2.0;
            if (d == 2) return -eta * (1 - zeta * zeta) * (1 + a * xi) * a * xi;
            return -zeta * (1 - eta * eta) * (1 + a * xi) * a * xi;
        case NC_FACEY:
            if (d == 1) return -xi * (1 - zeta * zeta) * (1 + b * eta) * b * eta;
            if (d == 2) return (1 - xi * xi) * (1 - zeta * zeta) * (1 + 2 * b * eta) * b / 2.0;
            return -zeta * (1 - xi * xi) * (1 + b * eta) * b * eta;
        case NC_FACEZ:
            if (d == 1) return -xi * (1 - eta * eta) * (1 + c * zeta) * c * zeta;
            if (d == 2) return -eta * (1 - xi * xi) * (1 + c * zeta) * c * zeta;
            return (1 - xi * xi) * (1 - eta * eta) * (1 + 2 * c * zeta) * c / 2.0;
        default:
            if (d == 1) return -2.0 * xi * (1 - eta * eta) * (1 - zeta * zeta);
            if (d == 2) return -2.0 * eta * (1 - xi * xi) * (1 - zeta * zeta);
            return -2.0 * zeta * (1 - xi * xi) * (1 - eta * eta);
        }
    }

    // v_fem.f90:78-179; i 1-based node, dir 1..3.  Combinations the reference leaves
    // undefined (function result never assigned) are never requested by mx_edge_dir.
    double mix_ln(int i, int dir, double xi, double eta, double zeta) const {
        const double a = nr[i - 1][0], b = nr[i - 1][1], c = nr[i - 1][2];
        if (mn == 8) {
            if (dir == 1) return ((1 + b * eta) * (1 + c * zeta)) / 4.0;
            if (dir == 2) return ((1 + a * xi) * (1 + c * zeta)) / 4.0;
            return ((1 + a * xi) * (1 + b * eta)) / 4.0;
        }
        if (mn == 20) {
            switch (node_class(i)) {
            case NC_CORNER:
                if (dir == 1) return ((1 + b * eta) * (1 + c * zeta) * (a * xi + b * eta + c * zeta - 1.0)) / 8.0;
                if (dir == 2) return ((1 + a * xi) * (1 + c * zeta) * (a * xi + b * eta + c * zeta - 1)) / 8.0;
                return ((1 + a * xi) * (1 + b * eta) * (a * xi + b * eta + c * zeta - 1)) / 8.0;
            case NC_MIDX:
                if (dir == 2) return ((1 - xi * xi) * (1 + c * zeta)) / 2.0;
                return ((1 - xi * xi) * (1 + b * eta)) / 2.0;
            case NC_MIDY:
                if (dir == 1) return ((1 - eta * eta) * (1 + c * zeta)) / 2.0;
                return ((1 + a * xi) * (1 - eta * eta)) / 2.0;
            case NC_MIDZ:
                if (dir == 1) return ((1 + b * eta) * (1 - zeta * zeta)) / 2.0;
                return ((1 + a * xi) * (1 - zeta * zeta)) / 2.0;
            default: return 0.0;
            }
        }
        switch (node_class(i)) {
        case NC_CORNER:
            if (dir == 1) return ((1 + a * xi) * (1 + b * eta) * (1 + c * zeta) * (b * eta * c * zeta)) / 8.0;
            if (dir == 2) return ((1 + a * xi) * (1 + b * eta) * (1 + c * zeta) * (a * xi * c * zeta)) / 8.0;
            return ((1 + a * xi) * (1 + b * eta) * (1 + c * zeta) * (a * xi * b * eta)) / 8.0;
        case NC_MIDX:
            if (dir == 2) return ((1 - xi * xi) * (1 + b * eta) * (1 + c * zeta) * c * zeta) / 4.0;
            return ((1 - xi * xi) * (1 + b * eta) * (1 + c * zeta) * b * eta) / 4.0;
        case NC_MIDY:
            if (dir == 1) return ((1 + a * xi) * (1 - eta * eta) * (1 + c * zeta) * c * zeta) / 4.0;
            return ((1 + a * xi) * (1 - eta * eta) * (1 + c * zeta) * a * xi) / 4.0;
        case NC_MIDZ:
            if (dir == 1) return ((1 + a * xi) * (1 + b * eta) * (1 - zeta * zeta) * b * eta) / 4.0;
            return ((1 + a * xi) * (1 + b * eta) * (1 - zeta * zeta) * a * xi) / 4.0;
        case NC_FACEX: return (1 - eta * eta) * (1 - zeta * zeta) * (1 + a * xi) / 2.0;
        case NC_FACEY: return (1 - xi * xi) * (1 - zeta * zeta) * (1 + b * eta) / 2.0;
        case NC_FACEZ: return (1 - xi * xi) * (1 - eta * eta) * (1 + c * zeta) / 2.0;
        default: return 0.0;
        }
    }

    // v_fem.f90:184-466; dir = basis direction, d = derivative axis, i 1-based node
    double mix_dln_dxi(int dir, int d, int i, double xi, double eta, double zeta) const {
        const double a = nr[i - 1][0], b = nr[i - 1][1], c = nr[i - 1][2];
        if (mn == 8) {
            if (dir == 1) {
                if (d == 1) return 0.0;
                if (d == 2) return b * (1 + c * zeta) / 4.0;
                return c * (1 + b * eta) / 4.0;
            }
            if (dir == 2) {
                if (d == 1) return a * (1 + c * zeta) / 4.0;
                if (d == 2) return 0.0;
                return c * (1 + a * xi) / 4.0;
            }
            if (d == 1) return a * (1 + b * eta) / 4.0;
            if (d == 2) return b * (1 + a * xi) / 4.0;
            return 0.0;
        }
        if (mn == 20) {
            switch (node_class(i)) {
            case NC_CORNER:
                if (dir == 1) {
                    if (d == 1) return ((1 + b * eta) * (1 + c * zeta) * (a)) / 8.0;
                    if (d == 2) return ((b) * (1 + c * zeta) * (a * xi + 2 * b * eta + c * zeta)) / 8.0;
                    return ((1 + b * eta) * (c) * (a * xi + b * eta + 2 * c * zeta)) / 8.0;
                }
                if (dir == 2) {
                    if (d == 1) return ((a) * (1 + c * zeta) * (2 * a * xi + b * eta + c * zeta)) / 8.0;
                    if (d == 2) return ((1 + a * xi) * (1 + c * zeta) * (b)) / 8.0;
                    return ((1 + a * xi) * (c) * (a * xi + b * eta + 2 * c * zeta)) / 8.0;
                }
                if (d == 1) return ((a) * (1 + b * eta) * (2 * a * xi + b * eta + c * zeta)) / 8.0;
                if (d == 2) return ((1 + a * xi) * (b) * (a * xi + 2 * b * eta + c * zeta)) / 8.0;
                return ((1 + a * xi) * (1 + b * eta) * (c)) / 8.0;
            case NC_MIDX:
                if (dir == 2) {
                    if (d == 1) return ((-xi) * (1 + c * zeta));
                    if (d == 2) return 0.0;
                    return ((1 - xi * xi) * (c)) / 2.0;
                }
                if (d == 1) return ((-xi) * (1 + b * eta));
                if (d == 2) return ((1 - xi * xi) * (b)) / 2.0;
                return 0.0;
            case NC_MIDY:
                if (dir == 1) {
                    if (d == 1) return 0.0;
                    if (d == 2) return ((-eta) * (1 + c * zeta));
                    return ((1 - eta * eta) * (c)) / 2.0;
                }
                if (d == 1) return ((a) * (1 - eta * eta)) / 2.0;
                if (d == 2) return ((1 + a * xi) * (-eta));
                return 0.0;
            case NC_MIDZ:
                if (dir == 1) {
                    if (d == 1) return 0.0;
                    if (d == 2) return ((b) * (1 - zeta * zeta)) / 2.0;
                    return ((1 + b * eta) * (-zeta));
                }
                if (d == 1) return ((a) * (1 - zeta * zeta)) / 2.0;
                if (d == 2) return 0.0;
                return ((1 + a * xi) * (-zeta));
            default: return 0.0;
            }
        }
        switch (node_class(i)) {
        case NC_CORNER:
            if (dir == 1) {
                if (d == 1) return ((a) * (1 + b * eta) * (1 + c * zeta) * (b * eta * c * zeta)) / 8.0;
                if (d == 2) return ((1 + a * xi) * (1 + 2 * b * eta) * (1 + c * zeta) * (b * c * zeta)) / 8.0;
                return ((1 + a * xi) * (1 + b * eta) * (1 + 2 * c * zeta) * (b * eta * c)) / 8.0;
            }
            if (dir == 2) {
                if (d == 1) return ((1 + 2 * a * xi) * (1 + b * eta) * (1 + c * zeta) * (a * c * zeta)) / 8.0;
                if (d == 2) return ((1 + a * xi) * (b) * (1 + c * zeta) * (a * xi * c * zeta)) / 8.0;
                return ((1 + a * xi) * (1 + b * eta) * (1 + 2 * c * zeta) * (a * xi * c)) / 8.0;
            }
            if (d == 1) return ((1 + 2 * a * xi) * (1 + b * eta) * (1 + c * zeta) * (a * b * eta)) / 8.0;
            if (d == 2) return ((1 + a * xi) * (1 + 2 * b * eta) * (1 + c * zeta) * (a * xi * b)) / 8.0;
            return ((1 + a * xi) * (1 + b * eta) * (c) * (a * xi * b * eta)) / 8.0;
        case NC_MIDX:
            if (dir == 2) {
                if (d == 1) return ((-xi) * (1 + b * eta) * (1 + c * zeta) * c * zeta) / 2.0;
                if (d == 2) return ((1 - xi * xi) * (b) * (1 + c * zeta) * c * zeta) / 4.0;
                return ((1 - xi * xi) * (1 + b * eta) * (1 + 2 * c * zeta) * c) / 4.0;
            }
            if (d == 1) return ((-xi) * (1 + b * eta) * (1 + c * zeta) * b * eta) / 2.0;
            if (d == 2) return ((1 - xi * xi) * (1 + 2 * b * eta) * (1 + c * zeta) * b) / 4.0;
            return ((1 - xi * xi) * (1 + b * eta) * (c) * b * eta) / 4.0;
        case NC_MIDY:
            if (dir == 1) {
                if (d == 1) return ((a) * (1 - eta * eta) * (1 + c * zeta) * c * zeta) / 4.0;
                if (d == 2) return ((1 + a * xi) * (-eta) * (1 + c * zeta) * c * zeta) / 2.0;
                return ((1 + a * xi) * (1 - eta * eta) * (1 + 2 * c * zeta) * c) / 4.0;
            }
            if (d == 1) return ((1 + 2 * a * xi) * (1 - eta * eta) * (1 + c * zeta) * a) / 4.0;
            if (d == 2) return ((1 + a * xi) * (-eta) * (1 + c * zeta) * a * xi) / 2.0;
            return ((1 + a * xi) * (1 - eta * eta) * (c) * a * xi) / 4.0;
        case NC_MIDZ:
            if (dir == 1) {
                if (d == 1) return ((a) * (1 + b * eta) * (1 - zeta * zeta) * b * eta) / 4.0;
                if (d == 2) return ((1 + a * xi) * (1 + 2 * b * eta) * (1 - zeta * zeta) * b) / 4.0;
                return ((1 + a * xi) * (1 + b * eta) * (-zeta) * b * eta) / 2.0;
            }
            if (d == 1) return ((1 + 2 * a * xi) * (1 + b * eta) * (1 - zeta * zeta) * a) / 4.0;
            if (d == 2) return ((1 + a * xi) * (b) * (1 - zeta * zeta) * a * xi) / 4.0;
            return ((1 + a * xi) * (1 + b * eta) * (-zeta) * a * xi) / 2.0;
        case NC_FACEX:
            if (d == 1) return (1 - eta * eta) * (1 - zeta * zeta) * (a) / 2.0;
            if (d == 2) return (-eta) * (1 - zeta * zeta) * (1 + a * xi);
            return (1 - eta * eta) * (-zeta) * (1 + a * xi);
        case NC_FACEY:
            if (d == 1) return (-xi) * (1 - zeta * zeta) * (1 + b * eta);
            if (d == 2) return (1 - xi * xi) * (1 - zeta * zeta) * (b) / 2.0;
            return (1 - xi * xi) * (-zeta) * (1 + b * eta);
        case NC_FACEZ:
            if (d == 1) return (-xi) * (1 - eta * eta) * (1 + c * zeta);
            if (d == 2) return (1 - xi * xi) * (-eta) * (1 + c * zeta);
            return (1 - xi * xi) * (1 - eta * eta) * (c) / 2.0;
        default: return 0.0;
        }
    }
};

// ---- integration.f90:297-299,329-331,361-363 / 400-402: DEFAULT-REAL literals (Q1) -------
inline int gauss_rule(int me, double *pt, double *wt) {
    if (me == 12) {
        pt[0] = -(double)0.577350269189625764509148780502f; pt[1] = (double)0.577350269189625764509148780502f;
        wt[0] = 1.0; wt[1] = 1.0;                                   // 1.d0 literals are exact
        return 2;
    }
    pt[0] = -(double)0.774596669241483377035853079956f; pt[1] = (double)0.0f;
    pt[2] = (double)0.774596669241483377035853079956f;
    wt[0] = (double)0.555555555555555555555555555556f; wt[1] = (double)0.888888888888888888888888888889f;
    wt[2] = (double)0.555555555555555555555555555556f;
    return 3;
}

}  // namespace movfem
#endif

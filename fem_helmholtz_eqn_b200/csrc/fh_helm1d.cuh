// Row generator of the reference's 1D P1 Helmholtz system, shared by the band-assembly kernel
// and the fused assemble+solve kernel so that both produce bit-identical coefficients.
//
// Every arithmetic step uses the explicitly rounded intrinsics (__dmul_rn / __dadd_rn / ...),
// which nvcc never contracts into FMAs, and follows the reference's evaluation order:
//   element_matrices   nb:70-96   jac = h/2; g = dphi/jac; Ke = ((g_i*g_j)*jac)*w; Me = ((phi_i*phi_j)*jac)*w
//   assembly scatter   nb:161-168 K[j,j] = 0 + Ke(e=j-1)[1,1] + Ke(e=j)[0,0]   (element order)
//   system matrix      nb:185     A = -K + (k**2)*M + S
//   boundary_matrices  nb:99-131  b[0] = -i*k*cos(theta);  A[N,N] += i*k
//   Dirichlet rows     solvers/fem2d_dirichlet.py:20-27   A[row]=0; A[row,row]=1; b[row]=g
#pragma once
#include "fh_common.cuh"

namespace fh {

struct Elem1d {  // the four distinct numbers of one element's Ke / Me pair
    double k00, k01, k10, k11, m00, m01, m10, m11;
};

__device__ __forceinline__ void shape1d(double xi, double phi[2], double dphi[2]) {  // nb:55-67
    phi[0] = __ddiv_rn(__dsub_rn(1.0, xi), 2.0);
    phi[1] = __ddiv_rn(__dadd_rn(1.0, xi), 2.0);
    dphi[0] = -0.5;
    dphi[1] = 0.5;
}

__device__ __forceinline__ Elem1d elem1d(double h, double xi, double w) {  // nb:70-96
    double phi[2], dphi[2];
    shape1d(xi, phi, dphi);
    const double jac = __ddiv_rn(h, 2.0);
    const double g0 = __ddiv_rn(dphi[0], jac), g1 = __ddiv_rn(dphi[1], jac);
    Elem1d e;
    e.k00 = __dmul_rn(__dmul_rn(__dmul_rn(g0, g0), jac), w);
    e.k01 = __dmul_rn(__dmul_rn(__dmul_rn(g0, g1), jac), w);
    e.k10 = __dmul_rn(__dmul_rn(__dmul_rn(g1, g0), jac), w);
    e.k11 = __dmul_rn(__dmul_rn(__dmul_rn(g1, g1), jac), w);
    e.m00 = __dmul_rn(__dmul_rn(__dmul_rn(phi[0], phi[0]), jac), w);
    e.m01 = __dmul_rn(__dmul_rn(__dmul_rn(phi[0], phi[1]), jac), w);
    e.m10 = __dmul_rn(__dmul_rn(__dmul_rn(phi[1], phi[0]), jac), w);
    e.m11 = __dmul_rn(__dmul_rn(__dmul_rn(phi[1], phi[1]), jac), w);
    return e;
}

// k-independent part of global row j: the element-order sums of K and M on the three bands.
struct RowGeom {
    double Kd, Md;  // diagonal:       K[j,j],   M[j,j]
    double Kl, Ml;  // sub-diagonal:   K[j,j-1], M[j,j-1]   (0 for j = 0)
    double Ku, Mu;  // super-diagonal: K[j,j+1], M[j,j+1]   (0 for j = N)
};

// Device copy of fh_problem1d (plain data, passed by value to kernels).
struct Problem1d {
    int64_t n_elem;
    double h;
    const double *x;
    int h_mode, bc_left, bc_right;
    double gl_re, gl_im, gr_re, gr_im, cos_theta, xi, w;
};

__device__ __forceinline__ double elem_size(const Problem1d &p, int64_t e) {
    return p.h_mode == FH_H_NODAL ? __dsub_rn(__ldg(p.x + e + 1), __ldg(p.x + e)) : p.h;
}

__device__ __forceinline__ RowGeom row_geom(const Problem1d &p, int64_t j) {
    RowGeom g = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    if (j >= 1) {  // element e = j-1 is scattered first (nb:161)
        const Elem1d L = elem1d(elem_size(p, j - 1), p.xi, p.w);
        g.Kd = __dadd_rn(g.Kd, L.k11);
        g.Md = __dadd_rn(g.Md, L.m11);
        g.Kl = L.k10;
        g.Ml = L.m10;
    }
    if (j < p.n_elem) {  // then element e = j
        const Elem1d R = elem1d(elem_size(p, j), p.xi, p.w);
        g.Kd = __dadd_rn(g.Kd, R.k00);
        g.Md = __dadd_rn(g.Md, R.m00);
        g.Ku = R.k01;
        g.Mu = R.m01;
    }
    return g;
}

// Row j of (A, b) for wavenumber k (k2 = the caller's k**2).
__device__ __forceinline__ void row_values(const Problem1d &p, const RowGeom &g, int64_t j, double k,
                                           double k2, cplx &dl, cplx &d, cplx &du, cplx &b) {
    const int64_t N = p.n_elem;
    dl = mk(j >= 1 ? __dadd_rn(-g.Kl, __dmul_rn(k2, g.Ml)) : 0.0, 0.0);
    du = mk(j < N ? __dadd_rn(-g.Ku, __dmul_rn(k2, g.Mu)) : 0.0, 0.0);
    d = mk(__dadd_rn(-g.Kd, __dmul_rn(k2, g.Md)), 0.0);
    b = czero();
    if (j == 0) b.im = -__dmul_rn(k, p.cos_theta);  // nb:113-121: Ne[0] = -i*k*cos(theta)
    if (j == N) d.im = k;                           // nb:124-129: Se[1,1] = i*k
    if (j == 0 && p.bc_left == FH_BC_DIRICHLET) {
        d = cone();
        du = czero();
        b = mk(p.gl_re, p.gl_im);
    }
    if (j == N && p.bc_right == FH_BC_DIRICHLET) {
        d = cone();
        dl = czero();
        b = mk(p.gr_re, p.gr_im);
    }
}

inline int problem_from_host(const fh_problem1d *h, Problem1d &p) {
    FH_REQUIRE(h != nullptr, "problem descriptor is NULL");
    FH_REQUIRE(h->n_elem >= 1, "n_elem must be >= 1 (got %lld)", (long long)h->n_elem);
    FH_REQUIRE(h->h_mode == FH_H_SCALAR || h->h_mode == FH_H_NODAL, "bad h_mode %d", h->h_mode);
    FH_REQUIRE(h->h_mode == FH_H_SCALAR || h->x != nullptr, "FH_H_NODAL needs the node array x");
    FH_REQUIRE(h->h_mode == FH_H_NODAL || h->h > 0.0, "scalar h must be > 0");
    FH_REQUIRE(h->bc_left == FH_BC_NEUMANN || h->bc_left == FH_BC_DIRICHLET,
               "bc_left must be NEUMANN or DIRICHLET (got %d)", h->bc_left);
    FH_REQUIRE(h->bc_right == FH_BC_SOMMERFELD || h->bc_right == FH_BC_DIRICHLET,
               "bc_right must be SOMMERFELD or DIRICHLET (got %d)", h->bc_right);
    p.n_elem = h->n_elem;
    p.h = h->h;
    p.x = h->x;
    p.h_mode = h->h_mode;
    p.bc_left = h->bc_left;
    p.bc_right = h->bc_right;
    p.gl_re = h->g_left_re;
    p.gl_im = h->g_left_im;
    p.gr_re = h->g_right_re;
    p.gr_im = h->g_right_im;
    p.cos_theta = h->cos_theta;
    p.xi = h->gauss_xi;
    p.w = h->gauss_w;
    return FH_OK;
}

}  // namespace fh

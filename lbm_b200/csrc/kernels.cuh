// kernels.cuh -- sm_100a device code of the lattice-Boltzmann sweep.
//
//   K1  sweep_kernel        fused pull-stream + BGK + link-wise boundaries
//                           (Domain::stream domain.hpp:116-141, Domain::collide
//                           :144-166, BGKCollision::collide collision.hpp:61-70,
//                           boundary.hpp:15-214)
//   K1g ghost_fluid_kernel  in-place BGK of ghost-shell cells left "fluid"
//                           (domain.hpp:147-155 loops 0..l+1)
//   K2  materialize_kernel  reference-style boundary pass (domain.hpp:157-165),
//                           run only before populations are read back
//   K3  macroscopic_kernel  density / velocity read-out (io/vtk.hpp:62-73)
//   K5  init / layout / mask kernels
//
// Data layout (HBM): structure of arrays  f[buffer][q][z][y][x]  with rows padded
// to P = roundup(xl+2,16) doubles and shifted by 15 elements so that interior
// x = 1 starts a 128-byte line.  A row's last ghost element (x = xl+1) spills
// into element 0 of the next row's padding, which that row never uses.
#pragma once
#include <cstdint>
#include <cuda.h>      // CUtensorMap (type only; the encoder is fetched through cudaGetDriverEntryPoint)
#include "lattice.cuh"

namespace lbmb200 {

enum : int {
    K_FLUID = 0, K_NOSLIP = 1, K_MOVINGWALL = 2, K_FREESLIP = 3, K_OUTFLOW = 4,
    K_INFLOW = 5, K_PRESSURE = 6, K_NULL = 7, K_PARALLEL = 8, K_PERIODIC = 9, K_COUNT = 10
};

constexpr uint32_t MASK_SKIP = 0x80000000u;       // not streamed: not an interior cell, or not fluid in the source lattice
constexpr uint32_t MASK_NOCOLLIDE = 0x40000000u;  // streamed, but the destination lattice's handler is not the fluid one
                                                  // (only when the two lattices carry different handlers, see GeoLayer)
constexpr uint32_t MASK_ALLNOSLIP = 0x20000000u;  // every flagged pull source is a NoSlipBoundary cell: the wall path needs
                                                  // no handler look-ups, only the cell's own inverse populations
constexpr uint32_t MASK_ALLPERIODIC = 0x10000000u; // every flagged pull source is a PERIODIC ghost cell: wrapped pulls
constexpr uint32_t MASK_ONEHANDLER = 0x08000000u;  // every flagged pull source carries the SAME handler object (kind and id):
                                                   // looked up once (an inflow / outflow / moving-wall face)
constexpr int X_SHIFT = 15;                   // element offset of x = 0 inside a row
constexpr int TMA_X0 = X_SHIFT - 1;           // the tensor maps of the TMA-fed sweep view rows from this element on:
                                              // 16-byte aligned, x = 0 .. xl+1 at coordinates 1 .. xl+2 of ONE row of
                                              // pitch P (cells x > P-16 of a row live in the next row's padding)

struct BcRec {          // one boundary handler object, device copy
    int kind;
    int pad;
    double v[3];
    double rho;
    double feq[27];     // InflowBoundary: compute_feq(rho, v), precomputed on the host
};

struct Layout {
    int xl, yl, zl;     // local interior lengths (of a slab: the planes it owns along the split axis)
    int P;              // row pitch (elements)
    int plane;          // elements of one plane of the SLOW (= split) axis: P * (other axis + 2)
    int sy, sz;         // element strides of y and z: (P, plane) for z-slabs -- z slowest, like domain.hpp:61-64 --
                        // and (plane, P) for y-slabs, whose x-z planes are the contiguous ones
    int swap;           // 1: y is the slow axis
    long long qstride;  // elements between consecutive q arrays
};

LBM_HD inline int cell_at(const Layout& g, int x, int y, int z)
{
    return z * g.sz + y * g.sy + x + X_SHIFT;
}
LBM_HD inline int n_slow(const Layout& g) { return g.swap ? g.yl : g.zl; }
LBM_HD inline int n_mid(const Layout& g) { return g.swap ? g.zl : g.yl; }

struct SweepParams {
    const double* __restrict__ src;   // collide field of the previous step
    double* __restrict__ dst;         // becomes the collide field
    const uint32_t* __restrict__ mask;
    const uint32_t* __restrict__ bits;   // 1 bit per cell: mask != 0 (bulk cells never read `mask`)
    const uint8_t* __restrict__ kind;
    const uint16_t* __restrict__ bcid;
    const BcRec* __restrict__ bc;
    Layout g;
    int z0;            // first plane (of the slow axis) swept by this launch
    int z_step;        // plane stride between consecutive blockIdx.z (1, or zl-1 for the two edge planes)
    int bx_shift;      // log2(threads along x per block)
    int first;         // 1: boundary cells still hold host-visible values -> pull stored
    int xhint_lo, xhint_hi;   // what most cells of the x = 1 / x = xl face see in the ghost plane next to them (0: unknown,
                       // 1: NoSlipBoundary, 2: PERIODIC): their pulls out of that plane are pointed at the values
                       // the wall path would fetch, and kept if the link mask then confirms the guess
    int wrap_z;        // the periodic slow axis is closed inside this slab
    double tau;
    double omega;      // 1/tau (fast mode)
    // per-population base pointers with the pull offset folded in, so that a bulk
    // load is ONE 64-bit multiply-add on a constant-bank operand:
    //   srcq[q] = src + q*qstride - (cz*plane + cy*P + cx),   dstq[q] = dst + q*qstride
    const double* srcq[27];
    double* dstq[27];
    // optional remote copies of the slab-edge populations (peer ghost planes)
    double* up_dst;    // receives the c_slow=+1 populations of the last plane of the slow axis
    double* dn_dst;    // receives the c_slow=-1 populations of its first plane
    long long up_qstride, dn_qstride;
    long long up_off, dn_off;   // element offset of the target ghost plane
};

// ---------------------------------------------------------------------------
// small runtime tables (slow paths only)
template <int Q>
struct Tables {
    signed char c[27][3];
    signed char q_of_cube[27];
    double w[27];
};
template <int Q>
constexpr Tables<Q> make_tables()
{
    Tables<Q> t{};
    for (int i = 0; i < 27; ++i) t.q_of_cube[i] = -1;
    for (int q = 0; q < Q; ++q) {
        t.c[q][0] = (signed char) Lattice<Q>::cx(q);
        t.c[q][1] = (signed char) Lattice<Q>::cy(q);
        t.c[q][2] = (signed char) Lattice<Q>::cz(q);
        t.q_of_cube[Lattice<Q>::cube(q)] = (signed char) q;
        t.w[q] = Lattice<Q>::w(q);
    }
    return t;
}
__constant__ Tables<15> g_tab15 = make_tables<15>();
__constant__ Tables<19> g_tab19 = make_tables<19>();
__constant__ Tables<27> g_tab27 = make_tables<27>();
template <int Q> __device__ __forceinline__ const Tables<Q>& tables();
template <> __device__ __forceinline__ const Tables<15>& tables<15>() { return g_tab15; }
template <> __device__ __forceinline__ const Tables<19>& tables<19>() { return g_tab19; }
template <> __device__ __forceinline__ const Tables<27>& tables<27>() { return g_tab27; }

// ---------------------------------------------------------------------------
// arithmetic.  EXACT: IEEE add/mul/div in the reference's association, never
// contracted (the __d*_rn intrinsics are not fused by nvcc).
template <bool EXACT> struct Ar;
template <> struct Ar<true> {
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
};

// Correctly rounded quotient a / b from y = RN(1 / b), without a division: the reference divides by the same
// few values all the time (C_S*C_S, 2*C_S^4, 2*C_S^2, tau: constants; rho: three times per cell), 61 / 85
// IEEE divisions per D3Q19 / D3Q27 cell, which made the bit-identical mode 3x slower than the fast one.
//   q0 = RN(a*y)          |q0 - a/b| < 1.5 ulp   (y carries a relative error < 2^-53)
//   q1 = RN(q0 + r0*y),   r0 = RN(a - b*q0)      -> q1 is a FAITHFUL rounding of a/b (error of r0*y << 1 ulp)
//   q2 = RN(q1 + r1*y),   r1 = a - b*q1 (exact)  -> q2 = RN(a/b): Markstein's theorem (IBM J. R&D 34, 1990: q faithful,
//                                                  y within 2^-53 relative of 1/b, residual by FMA => correct rounding)
// Valid while the residuals neither underflow nor overflow (|a| and |a/b| within 2^+-900) and a is not -0:
// every dividend on this path is +0 or a sum / product of O(1e-18 .. 1) numbers.  Non-finite states (a run
// that has blown up) may propagate NaN where the reference still holds inf.  tools/selftest/div_check.c and
// lbm_b200_selftest_division compare the sequence with IEEE division on 10^8 operands.
#ifndef LBM_EXACT_DIV          // 0: IEEE divisions (__ddiv_rn) as in round 1, 1: the sequence above
#define LBM_EXACT_DIV 1
#endif
#ifndef LBM_DIV_STEPS          // 2 = proven; 1 = one correction only (measurement knob, NOT proven for q0 off by > 1 ulp)
#define LBM_DIV_STEPS 2
#endif
__device__ __forceinline__ double div_rcp(double a, double b, double y)
{
#if LBM_EXACT_DIV
    double q = __dmul_rn(a, y);
    double r = __fma_rn(-b, q, a);
    q = __fma_rn(r, y, q);
#if LBM_DIV_STEPS >= 2
    r = __fma_rn(-b, q, a);
    q = __fma_rn(r, y, q);
#endif
    return q;
#else
    return __ddiv_rn(a, b);
#endif
}
constexpr double RCP_CS2 = 1.0 / CS2, RCP_TWO_CS4 = 1.0 / TWO_CS4, RCP_TWO_CS2 = 1.0 / TWO_CS2;   // correctly rounded
__device__ __forceinline__ double div_cs2(double a) { return div_rcp(a, CS2, RCP_CS2); }
__device__ __forceinline__ double div_two_cs4(double a) { return div_rcp(a, TWO_CS4, RCP_TWO_CS4); }
__device__ __forceinline__ double div_two_cs2(double a) { return div_rcp(a, TWO_CS2, RCP_TWO_CS2); }
// 1 / b for a run-time divisor used several times (rho): one correctly rounded reciprocal
__device__ __forceinline__ double rcp_exact(double b)
{
#if LBM_EXACT_DIV
    return __drcp_rn(b);
#else
    return 0.0;
#endif
}

// c_dot_u of collision.hpp:40-44, ((0 + c0*u0) + c1*u1) + c2*u2.  Terms with c = 0 are skipped and c = +-1 is
// an add / subtract: +-1 * u is exact, 0 * u = +-0 leaves a sum that is never -0 untouched (the chain starts
// from +0, and x + (-x) = +0), so the bits are the reference's for all finite u.
template <int Q, int q>
__device__ __forceinline__ double cu_exact(double ux, double uy, double uz)
{
    using L = Lattice<Q>;
    using A = Ar<true>;
    double cu = 0.0;
    if constexpr (L::cx(q) == 1) cu = A::add(cu, ux);
    if constexpr (L::cx(q) == -1) cu = A::sub(cu, ux);
    if constexpr (L::cy(q) == 1) cu = A::add(cu, uy);
    if constexpr (L::cy(q) == -1) cu = A::sub(cu, uy);
    if constexpr (L::cz(q) == 1) cu = A::add(cu, uz);
    if constexpr (L::cz(q) == -1) cu = A::sub(cu, uz);
    return cu;
}

// collision.hpp:34-51, one entry; uu_term = u_dot_u / (2*C_S*C_S) is the same
// value for every q in the reference and is hoisted by the callers.
template <int Q, int q>
__device__ __forceinline__ double feq_exact(double rho, double ux, double uy, double uz, double uu_term)
{
    using L = Lattice<Q>;
    using A = Ar<true>;
    const double cu = cu_exact<Q, q>(ux, uy, uz);
    double t = A::add(1.0, div_cs2(cu));
    t = A::add(t, div_two_cs4(A::mul(cu, cu)));
    t = A::sub(t, uu_term);
    return A::mul(A::mul(L::w(q), rho), t);
}
// feq of direction q and of its inverse at once: c_dot_u of the inverse is -c_dot_u bit for bit (or both +0),
// a quotient changes sign with its dividend, 1 + (-d) = 1 - d, and the square is shared.
template <int Q, int q>
__device__ __forceinline__ void feq_pair_exact(double rho, double ux, double uy, double uz, double uu_term,
                                               double& e, double& e_inv)
{
    using L = Lattice<Q>;
    using A = Ar<true>;
    static_assert(L::w(q) == L::w(L::inv(q)), "opposite directions share their weight");
    const double cu = cu_exact<Q, q>(ux, uy, uz);
    const double d1 = div_cs2(cu);
    const double d2 = div_two_cs4(A::mul(cu, cu));
    const double wr = A::mul(L::w(q), rho);
    e = A::mul(wr, A::sub(A::add(A::add(1.0, d1), d2), uu_term));
    e_inv = A::mul(wr, A::sub(A::add(A::sub(1.0, d1), d2), uu_term));
}
__device__ __forceinline__ double uu_term_exact(double ux, double uy, double uz)
{
    using A = Ar<true>;
    double uu = A::add(0.0, A::mul(ux, ux));
    uu = A::add(uu, A::mul(uy, uy));
    uu = A::add(uu, A::mul(uz, uz));
    return div_two_cs2(uu);
}

// one term of compute_density / compute_velocity (collision.hpp:7-31): rho += v; m += v * c.  c = 0 terms add
// +-0 to sums that are never -0 and are skipped, c = +-1 is an add / subtract -- same bits for finite v.
template <int Q, int q>
__device__ __forceinline__ void moment_term_exact(double v, double& rho, double& mx, double& my, double& mz)
{
    using L = Lattice<Q>;
    using A = Ar<true>;
    rho = A::add(rho, v);
    if constexpr (L::cx(q) == 1) mx = A::add(mx, v);
    if constexpr (L::cx(q) == -1) mx = A::sub(mx, v);
    if constexpr (L::cy(q) == 1) my = A::add(my, v);
    if constexpr (L::cy(q) == -1) my = A::sub(my, v);
    if constexpr (L::cz(q) == 1) mz = A::add(mz, v);
    if constexpr (L::cz(q) == -1) mz = A::sub(mz, v);
}

// moments of a register-resident cell: collision.hpp:7-31
template <int Q>
__device__ __forceinline__ void moments_exact(const double (&f)[Q], double& rho, double& mx, double& my, double& mz)
{
    rho = 0.0; mx = 0.0; my = 0.0; mz = 0.0;
    static_for<Q>([&](auto I) {
        constexpr int q = decltype(I)::value;
        moment_term_exact<Q, q>(f[q], rho, mx, my, mz);
    });
}

// u = m / rho (collision.hpp:27-29): three quotients by the same divisor
__device__ __forceinline__ void velocity_exact(double rho, double mx, double my, double mz, double& ux, double& uy, double& uz)
{
    const double ir = rcp_exact(rho);
    ux = div_rcp(mx, rho, ir);
    uy = div_rcp(my, rho, ir);
    uz = div_rcp(mz, rho, ir);
}

template <int Q, bool EXACT>
__device__ __forceinline__ void bgk_collide(double (&f)[Q], double tau, double omega)
{
    using L = Lattice<Q>;
    if constexpr (EXACT) {
        using A = Ar<true>;
        double rho, mx, my, mz, ux, uy, uz;
        moments_exact<Q>(f, rho, mx, my, mz);
        velocity_exact(rho, mx, my, mz, ux, uy, uz);
        const double uut = uu_term_exact(ux, uy, uz);
        // f -= (f - feq) / tau, collision.hpp:68; omega = RN(1 / tau) from the host
        static_for<(Q + 1) / 2>([&](auto I) {
            constexpr int q = decltype(I)::value;
            constexpr int qi = L::inv(q);
            if constexpr (q == qi) {
                const double e = feq_exact<Q, q>(rho, ux, uy, uz, uut);
                f[q] = A::sub(f[q], div_rcp(A::sub(f[q], e), tau, omega));
            } else {
                double e, ei;
                feq_pair_exact<Q, q>(rho, ux, uy, uz, uut, e, ei);
                f[q] = A::sub(f[q], div_rcp(A::sub(f[q], e), tau, omega));
                f[qi] = A::sub(f[qi], div_rcp(A::sub(f[qi], ei), tau, omega));
            }
        });
    } else {
        // same formula, reciprocals and free contraction; opposite directions
        // share the even part of the equilibrium
        constexpr double I_CS2 = 1.0 / CS2, I_2CS4 = 1.0 / TWO_CS4, I_2CS2 = 1.0 / TWO_CS2;
        double rho = 0.0, mx = 0.0, my = 0.0, mz = 0.0;
        static_for<Q>([&](auto I) {
            constexpr int q = decltype(I)::value;
            rho += f[q];
            if constexpr (L::cx(q) == 1) mx += f[q];
            if constexpr (L::cx(q) == -1) mx -= f[q];
            if constexpr (L::cy(q) == 1) my += f[q];
            if constexpr (L::cy(q) == -1) my -= f[q];
            if constexpr (L::cz(q) == 1) mz += f[q];
            if constexpr (L::cz(q) == -1) mz -= f[q];
        });
        const double ir = 1.0 / rho;
        const double ux = mx * ir, uy = my * ir, uz = mz * ir;
        const double base = 1.0 - (ux * ux + uy * uy + uz * uz) * I_2CS2;
        static_for<(Q + 1) / 2>([&](auto I) {
            constexpr int q = decltype(I)::value;
            constexpr int qi = L::inv(q);
            const double wr = L::w(q) * rho;
            if constexpr (q == qi) {             // rest population
                f[q] -= (f[q] - wr * base) * omega;
            } else {
                const double cu = (double) L::cx(q) * ux + (double) L::cy(q) * uy + (double) L::cz(q) * uz;
                const double even = wr * (base + cu * cu * I_2CS4);
                const double odd = wr * (cu * I_CS2);
                f[q] -= (f[q] - (even + odd)) * omega;
                f[qi] -= (f[qi] - (even - odd)) * omega;
            }
        });
    }
}

// ---------------------------------------------------------------------------
// Link-wise boundary values.  The entry q of boundary cell
// B = X - c_q is only ever read by the fluid cell X, and the reference computes
// it from X's post-collision state of the previous step (boundary.hpp:26-28),
// i.e. from  src[.][X].
struct OwnMoments {
    double rho, mx, my, mz;
    bool have;
    // OutflowBoundary / PressureBoundary: velocity = momentum / REFERENCE density of handler `rec_of_u`
    // (boundary.hpp:144, 209) and its u.u term, the same for every link of one handler
    double ux, uy, uz, uut;
    const void* rec_of_u;
};

// sequential moments of the cell at index i of `src` (collision.hpp:7-31); the sums keep the reference's order.
// Out of line on purpose: at the call sites the Q pulled values are live, and the callee gets by with few registers
// (ptxas then loads and adds one population after the other whatever the batch size -- measured: batches of 1, 9 and
// Q within 1 % of each other on the channel config, profiles/variants_r09_exact_wall.txt).
#ifndef LBM_MOM_BATCH
#define LBM_MOM_BATCH 9
#endif
#ifndef LBM_MOM_INLINE         // 1: single-handler wall cells sum their own moments inline, all Q loads at once
#define LBM_MOM_INLINE 1
#endif
template <int Q, int B>
__device__ __forceinline__ void sum_moments(const double* __restrict__ src, long long qstride, int i, OwnMoments* m)
{
    double rho = 0.0, mx = 0.0, my = 0.0, mz = 0.0;
    static_for<(Q + B - 1) / B>([&](auto C) {
        constexpr int q0 = decltype(C)::value * B;
        constexpr int n = Q - q0 < B ? Q - q0 : B;
        double v[n];
        static_for<n>([&](auto J) { constexpr int j = decltype(J)::value; v[j] = src[(q0 + j) * qstride + i]; });
        static_for<n>([&](auto J) {
            constexpr int j = decltype(J)::value;
            moment_term_exact<Q, q0 + j>(v[j], rho, mx, my, mz);
        });
    });
    m->rho = rho; m->mx = mx; m->my = my; m->mz = mz; m->have = true;
}
template <int Q>
__device__ __noinline__ void load_moments(const double* __restrict__ src, long long qstride, int i, OwnMoments* m)
{
    sum_moments<Q, (LBM_MOM_BATCH < 1 ? Q : LBM_MOM_BATCH)>(src, qstride, i, m);
}

// FreeSlipBoundary::collide, boundary.hpp:98-112, for the link (B, q), B = X - c_q.
// Returns the value the reference leaves in B[q]; if no branch fires the stored
// value stays.
template <int Q>
__device__ __noinline__ double freeslip_value(const double* __restrict__ src, const uint8_t* __restrict__ kind,
                                              const Layout g, int iX, int q)
{
    const Tables<Q>& T = tables<Q>();
    const int dx = T.c[q][0], dy = T.c[q][1], dz = T.c[q][2];
    const int b = iX - (dz * g.sz + dy * g.sy + dx);
    auto at = [&](int ox, int oy, int oz) { return b + oz * g.sz + oy * g.sy + ox; };
    auto pick = [&](int cell, int u, int v, int w) {
        const int qq = T.q_of_cube[(w + 1) * 9 + (v + 1) * 3 + (u + 1)];
        return src[qq * g.qstride + cell];
    };
    int c;
    if (kind[c = at(dx, 0, 0)] == K_FLUID) return pick(c, -dx, dy, dz);
    if (kind[c = at(0, dy, 0)] == K_FLUID) return pick(c, dx, -dy, dz);
    if (kind[c = at(0, 0, dz)] == K_FLUID) return pick(c, dx, dy, -dz);
    if (kind[c = at(0, dy, dz)] == K_FLUID) return pick(c, dx, -dy, -dz);
    if (kind[c = at(dx, 0, dz)] == K_FLUID) return pick(c, -dx, dy, -dz);
    if (kind[c = at(dx, dy, 0)] == K_FLUID) return pick(c, -dx, -dy, dz);
    return src[q * g.qstride + b];
}

// value of B[q] for handler (k, rec), computed from fluid cell X (index iX) in `src`
template <int Q, bool EXACT, int q>
__device__ __forceinline__ double link_value(const double* __restrict__ src, const uint8_t* __restrict__ kind,
                                             const Layout& g, int iX, int k, const BcRec* __restrict__ rec,
                                             OwnMoments& om)
{
    using L = Lattice<Q>;
    constexpr int qi = L::inv(q);
    switch (k) {
    case K_NOSLIP:                                            // boundary.hpp:28
        return src[qi * g.qstride + iX];
    case K_MOVINGWALL: {                                      // boundary.hpp:57-65
        if (!om.have) load_moments<Q>(src, g.qstride, iX, &om);
        const double finv = src[qi * g.qstride + iX];
        constexpr double cx = L::cx(q), cy = L::cy(q), cz = L::cz(q);
        // c_dot_u = ((0 + c0*u0) + c1*u1) + c2*u2 ;  finv + (((2.0*w)*rho)*cu)/(C_S*C_S)
        double cu = __dadd_rn(0.0, __dmul_rn(cx, rec->v[0]));
        cu = __dadd_rn(cu, __dmul_rn(cy, rec->v[1]));
        cu = __dadd_rn(cu, __dmul_rn(cz, rec->v[2]));
        constexpr double two_w = 2.0 * L::w(q);
        if constexpr (EXACT) {
            return __dadd_rn(finv, div_cs2(__dmul_rn(__dmul_rn(two_w, om.rho), cu)));
        } else {
            return finv + two_w * om.rho * cu * (1.0 / CS2);
        }
    }
    case K_INFLOW:                                            // boundary.hpp:178
        return rec->feq[q];
    case K_OUTFLOW:                                           // boundary.hpp:143-147
    case K_PRESSURE: {                                        // boundary.hpp:208-211
        if (!om.have) load_moments<Q>(src, g.qstride, iX, &om);
        const double r = rec->rho;                            // momentum / REFERENCE density
        if (om.rec_of_u != (const void*) rec) {               // once per handler, not once per link
            velocity_exact(r, om.mx, om.my, om.mz, om.ux, om.uy, om.uz);
            om.uut = uu_term_exact(om.ux, om.uy, om.uz);
            om.rec_of_u = rec;
        }
        double e, ei;
        if constexpr (q < qi) feq_pair_exact<Q, q>(r, om.ux, om.uy, om.uz, om.uut, e, ei);
        else if constexpr (q > qi) feq_pair_exact<Q, qi>(r, om.ux, om.uy, om.uz, om.uut, ei, e);
        else e = ei = feq_exact<Q, q>(r, om.ux, om.uy, om.uz, om.uut);
        return __dsub_rn(__dadd_rn(e, ei), src[qi * g.qstride + iX]);
    }
    case K_FREESLIP:
        return freeslip_value<Q>(src, kind, g, iX, q);
    default: {                                                // NULL / PARALLEL: stored value
        return src[q * g.qstride + iX - (L::cz(q) * g.sz + L::cy(q) * g.sy + L::cx(q))];
    }
    }
}

__device__ __forceinline__ int wrap1(int v, int l) { return v < 1 ? v + l : (v > l ? v - l : v); }

// ---------------------------------------------------------------------------
// K1: one thread per interior cell.  Bulk cells: Q coalesced loads, BGK in registers,
// Q coalesced 128B-aligned stores -> 2*Q*8 bytes per update.
// tuning knobs (defaults = the measured best, see profiles/)
#ifndef LBM_SWEEP_THREADS
#define LBM_SWEEP_THREADS 64
#endif
// Resident 64-thread blocks per SM the register allocator must allow = the register budget of a launch.
// Until the end of round 2 the answer was "as many threads as possible" (16 / 16 / 12 blocks = 64 / 64 / 80 registers): the
// link mask was a dependent load and wall lanes were slow, so resident warps bought bandwidth.  With one memory round
// trip per cell for (nearly) every lane the pulls of 512-640 threads keep the DRAM busy, and what costs now are the one
// or two pulled values ptxas spills at 64 / 80 registers (their reload sits in front of the collision).  Sweep over the
// budget (cavity 512^3, fraction of the measured HBM rate; profiles/variants_r11_registers.txt):
//   D3Q15 fast   64: .979   72: .982   80: .987   96: .994 (natural count, 10 blocks)        exact  64: .982  80: .990  96: .977
//   D3Q19 fast   64: .958   72: .973   80: .976   88: .974   96: .982   116: .990 (8 blocks)  exact  64: .943  80: .981  96: .985  126: .968
//   D3Q27 fast   80: .933   96: .961   128: .968 (8 blocks)   148: .981 (6 blocks, x-face mode only)   exact  80: .909  96: .965  128: .973  160: .954
// The bit-checked modes (CHECKED / SPLIT: pipes, literal masks) keep the round-1 budgets: their pulls DO wait for a map.
#ifndef LBM_MB15
#define LBM_MB15 16
#endif
#ifndef LBM_MB19
#define LBM_MB19 16
#endif
#ifndef LBM_MB27
#define LBM_MB27 12
#endif
#ifndef LBM_MB15_FAST
#define LBM_MB15_FAST 9
#endif
#ifndef LBM_MB15_EXACT
#define LBM_MB15_EXACT 12
#endif
#ifndef LBM_MB19_FAST
#define LBM_MB19_FAST 8
#endif
#ifndef LBM_MB19_EXACT
#define LBM_MB19_EXACT 9
#endif
#ifndef LBM_MB27_FAST
#define LBM_MB27_FAST 8
#endif
#ifndef LBM_MB27_EXACT
#define LBM_MB27_EXACT 8
#endif
// D3Q27, fast arithmetic, x-face guesses in use (every lane is a one-round-trip lane): 6 blocks = 148 registers reach 0.981;
// the same budget without the guesses (channel: inflow / outflow lanes walk the wall path) loses 12 %, so only SWEEP_XFACE gets it
#ifndef LBM_MB27_FAST_XFACE
#define LBM_MB27_FAST_XFACE 6
#endif
template <int Q, bool EXACT, int MODE> struct MinBlocks {     // MODE: the SWEEP_* enum below (0 speculative, 3 x-face guesses)
    static constexpr bool speculative = MODE == 0 || MODE == 3;
    static constexpr int value = !speculative ? (Q == 15 ? LBM_MB15 : (Q == 19 ? LBM_MB19 : LBM_MB27))
                               : EXACT ? (Q == 15 ? LBM_MB15_EXACT : (Q == 19 ? LBM_MB19_EXACT : LBM_MB27_EXACT))
                                       : (Q == 15 ? LBM_MB15_FAST : (Q == 19 ? LBM_MB19_FAST : (MODE == 3 ? LBM_MB27_FAST_XFACE : LBM_MB27_FAST)));
};
#ifdef LBM_LOAD_CS
#define LBM_LD(ptr) __ldcs(ptr)
#else
#define LBM_LD(ptr) (*(ptr))
#endif
#ifdef LBM_STORE_CS
#define LBM_ST(ptr, v) __stcs(ptr, v)
#else
#define LBM_ST(ptr, v) (*(ptr) = (v))
#endif

// Everything after the pull: replace the directions whose source is not fluid (link-wise boundary
// values), collide, store, and hand slab-edge populations to the neighbour.  `f` holds the values
// pulled for ALL directions; `m` is the cell's link mask (0 for bulk cells).  ONE code path: a cell
// next to a wall patches f[] and falls through to the same collide + store instructions as a bulk
// cell (a separate wall-cell path or launch was measured and is slower, profiles/r02_sweep_modes.txt).
template <int Q, bool EXACT, bool SPLIT = false>
__device__ __forceinline__ void finish_cell(const SweepParams& p, double (&f)[Q], const uint32_t m,
                                            const int i, const int x, const int y, const int z)
{
    using L = Lattice<Q>;
    const Layout& g = p.g;
    if ((m & MASK_ALLNOSLIP) && !p.first) {
        // the common wall cell (plain bounce-back, boundary.hpp:28): B[q] = f_inv(q)(X) of the previous step for every
        // flagged direction -- independent loads of the cell's own populations, ONE round trip after the mask instead
        // of kind -> handler -> value per direction
        static_for<Q>([&](auto I) {
            constexpr int q = decltype(I)::value;
            if (m & (1u << q)) f[q] = p.src[L::inv(q) * g.qstride + i];
        });
    } else if (m & MASK_ALLPERIODIC) {
        // x and the middle axis always wrap inside the slab, the slow axis only if it is closed here
        static_for<Q>([&](auto I) {
            constexpr int q = decltype(I)::value;
            if (m & (1u << q)) {
                const int sx = wrap1(x - L::cx(q), g.xl);
                const int sy = (!g.swap || p.wrap_z) ? wrap1(y - L::cy(q), g.yl) : y - L::cy(q);
                const int sz = (g.swap || p.wrap_z) ? wrap1(z - L::cz(q), g.zl) : z - L::cz(q);
                f[q] = p.src[q * g.qstride + cell_at(g, sx, sy, sz)];
            }
        });
    } else if ((m & MASK_ONEHANDLER) && !p.first) {
        // one handler for all flagged directions: its kind and record are fetched once, through the first flagged source
        const Tables<Q>& T = tables<Q>();
        const int q0 = __ffs(m) - 1;
        const int s0 = i - (T.c[q0][2] * g.sz + T.c[q0][1] * g.sy + T.c[q0][0]);
        const int k = p.kind[s0];
        const BcRec* rec = p.bc + p.bcid[s0];
        OwnMoments om;
        om.have = false;
        om.rec_of_u = nullptr;
#if LBM_MOM_INLINE
        // with 96-148 registers the own populations of the cell fit next to the pulled ones: all Q loads of the moment sum
        // in flight at once, inlined here (the out-of-line version loads and adds one after the other)
        if (k == K_MOVINGWALL || k == K_OUTFLOW || k == K_PRESSURE) sum_moments<Q, Q>(p.src, g.qstride, i, &om);
#endif
        static_for<Q>([&](auto I) {
            constexpr int q = decltype(I)::value;
            if (m & (1u << q)) f[q] = link_value<Q, EXACT, q>(p.src, p.kind, g, i, k, rec, om);
        });
    } else if (m != 0) {
        OwnMoments om;
        om.have = false;
        om.rec_of_u = nullptr;
        static_for<Q>([&](auto I) {
            constexpr int q = decltype(I)::value;
            if (m & (1u << q)) {
                const int s = i - (L::cz(q) * g.sz + L::cy(q) * g.sy + L::cx(q));
                const int k = p.kind[s];
                if (k == K_PERIODIC) {
                    // x and the middle axis always wrap inside the slab, the slow axis only if it is closed here
                    const int sx = wrap1(x - L::cx(q), g.xl);
                    const int sy = (!g.swap || p.wrap_z) ? wrap1(y - L::cy(q), g.yl) : y - L::cy(q);
                    const int sz = (g.swap || p.wrap_z) ? wrap1(z - L::cz(q), g.zl) : z - L::cz(q);
                    f[q] = p.src[q * g.qstride + cell_at(g, sx, sy, sz)];
                } else if (!p.first) {   // first step: the stored value already pulled is the answer
                    f[q] = link_value<Q, EXACT, q>(p.src, p.kind, g, i, k, p.bc + p.bcid[s], om);
                }
            }
        });
    }
    if constexpr (SPLIT) {
        // SPLIT (the two lattices carry different handlers, "literal" edits): a cell can be streamed into although
        // its handler in the destination lattice is a boundary -- stored as streamed, not collided.  A separate
        // instantiation: carrying this exit in the common kernel costs D3Q27 1.8 % (profiles/variants_r08_wall.txt)
        if (m & MASK_NOCOLLIDE) {
            static_for<Q>([&](auto I) {
                constexpr int q = decltype(I)::value;
                p.dstq[q][i] = f[q];
            });
            return;
        }
    }

    bgk_collide<Q, EXACT>(f, p.tau, p.omega);

    static_for<Q>([&](auto I) {
        constexpr int q = decltype(I)::value;
        LBM_ST(p.dstq[q] + i, f[q]);
    });
    // slab edges: hand the populations that leave the slab to the neighbour (peer memory over NVLink)
    const int slow = g.swap ? y : z;
    if (p.up_dst != nullptr && slow == n_slow(g)) {
        const int ip = i - slow * g.plane;
        static_for<Q>([&](auto I) {
            constexpr int q = decltype(I)::value;
            if ((g.swap ? L::cy(q) : L::cz(q)) == 1) p.up_dst[q * p.up_qstride + p.up_off + ip] = f[q];
        });
    }
    if (p.dn_dst != nullptr && slow == 1) {
        const int ip = i - slow * g.plane;
        static_for<Q>([&](auto I) {
            constexpr int q = decltype(I)::value;
            if ((g.swap ? L::cy(q) : L::cz(q)) == -1) p.dn_dst[q * p.dn_qstride + p.dn_off + ip] = f[q];
        });
    }
}

// How a sweep launch finds out whether its cell is a bulk cell:
//   SWEEP_SPECULATIVE  the 1-bit map and the Q pulls are requested TOGETHER: every pull source of an interior
//                      cell exists in memory (ghost shell), so the loads need not wait for the map -> one DRAM
//                      round trip per cell instead of two dependent ones.  A cell that turns out not to be
//                      streamed (solid) has pulled Q values for nothing.
//   SWEEP_CHECKED      looks at the bit (L2-resident map) BEFORE pulling: a solid cell costs one map read instead
//                      of Q wasted pulls -- for geometries with large solid regions (pipe.vtk: 44 % solid).
//   SWEEP_SPLIT        SWEEP_CHECKED for lattices that carry different handlers (MASK_NOCOLLIDE cells exist)
//   SWEEP_XFACE        SWEEP_SPECULATIVE + guessed pulls for the cells of the two x faces (see sweep_kernel); chosen by the
//                      host when it has a guess (SweepParams::xhint_*).  Its own instantiation: the address selection
//                      costs a config without a guess 4-6 % (channel: inflow / outflow faces, profiles/variants_r10_xface.txt)
enum : int { SWEEP_SPECULATIVE = 0, SWEEP_CHECKED = 1, SWEEP_SPLIT = 2, SWEEP_XFACE = 3 };
static_assert(SWEEP_SPECULATIVE == 0 && SWEEP_XFACE == 3, "MinBlocks (above) names these two modes by value");

// bits of the directions with c_x = sign (the pulls of an x-face cell that leave the interior)
template <int Q>
constexpr uint32_t x_leaving_mask(int sign)
{
    uint32_t m = 0;
    for (int q = 0; q < Q; ++q)
        if (Lattice<Q>::cx(q) == sign) m |= 1u << q;
    return m;
}
#ifndef LBM_EAGER_MASK         // 1: cells of the six faces request their link mask together with the pulls
#define LBM_EAGER_MASK 1
#endif

template <int Q, bool EXACT, int MODE>
__global__ void __launch_bounds__(LBM_SWEEP_THREADS, MinBlocks<Q, EXACT, MODE>::value) sweep_kernel(const SweepParams p)
{
    using L = Lattice<Q>;
    constexpr bool SPECULATIVE = MODE == SWEEP_SPECULATIVE || MODE == SWEEP_XFACE;
    const Layout& g = p.g;
    const int bx = 1 << p.bx_shift;
    const int x = 1 + blockIdx.x * bx + (threadIdx.x & (bx - 1));
    const int mid = 1 + blockIdx.y * (LBM_SWEEP_THREADS >> p.bx_shift) + (threadIdx.x >> p.bx_shift);
    const int slow = p.z0 + blockIdx.z * p.z_step;          // planes of the slow axis: z, or y for y-slabs
    if (x > g.xl || mid > n_mid(g)) return;
    const int y = g.swap ? slow : mid, z = g.swap ? mid : slow;
    const int i = slow * g.plane + mid * g.P + x + X_SHIFT;
    // Bulk cells read 1/8 byte of map, not 4.
    const uint32_t word = p.bits[i >> 5];
    uint32_t m = 0;
    if constexpr (!SPECULATIVE) {
        if ((word >> (i & 31)) & 1u) {
            m = p.mask[i];
            if (m & MASK_SKIP) return;
        }
    }
    // SPECULATIVE.  A cell of one of the six faces is almost always a wall cell: its link mask is requested now,
    // together with the pulls, instead of after the bit map has arrived (one memory round trip less).
    // XFACE.  Every row has two x-face cells, i.e. a quarter (512 cells per row) or half (256) of all warps carry one
    // lane that walks the wall path -- mask, then the wall values, then collide -- while 31 lanes wait.  The host knows
    // what most of those cells see next to them (xhint); for a NoSlipBoundary the wall value of direction q is the
    // cell's own population inv(q) of the previous step (boundary.hpp:28), for a PERIODIC ghost plane it is the pull
    // from the opposite face: the lane simply pulls THAT address instead of the ghost cell's.  If the mask then says
    // exactly this (all flagged links are of the guessed kind and are the c_x = +-1 ones) the cell is done and joins
    // the bulk lanes; otherwise the guessed directions are pulled again from their true sources and the wall path runs.
    uint32_t m_eager = 0;
    bool face = false;
    int xs = 0, hint = 0;
    if constexpr (SPECULATIVE) {
        const bool yz_face = mid == 1 || mid == n_mid(g) || slow == 1 || slow == n_slow(g);
#if LBM_EAGER_MASK
        face = yz_face || x == 1 || x == g.xl;
        if (face) m_eager = p.mask[i];
#endif
        if constexpr (MODE == SWEEP_XFACE) {
            if (!yz_face) {        // (sources of an edge cell may wrap in y / z as well: left to the wall path)
                if (x == 1) { xs = 1; hint = p.xhint_lo; }
                else if (x == g.xl) { xs = -1; hint = p.xhint_hi; }
            }
        }
    }
    double f[Q];
    // Only the two warps of a row that hold an x-face lane select addresses (a warp-uniform branch; the selection
    // costs the loads of D3Q27 4-6 % when every warp carries it).  The guessed source of a c_x != 0 pull: bounce-back
    // -> the cell's own population inv(q), reached through the base pointer of inv(q) (which has its own pull offset
    // folded in: srcq[inv q] + i - c_q.strides = src[inv q][i]); periodic -> the same pull one period further along x.
    bool select = false;
    if constexpr (MODE == SWEEP_XFACE) select = __any_sync(__activemask(), hint != 0);
    if (select) {
        static_for<Q>([&](auto I) {
            constexpr int q = decltype(I)::value;
            if constexpr (L::cx(q) == 0) {
                f[q] = LBM_LD(p.srcq[q] + i);
            } else {
                const bool mine = xs == L::cx(q);
                if (mine && hint == 1) f[q] = LBM_LD(p.srcq[L::inv(q)] + (i - (L::cz(q) * g.sz + L::cy(q) * g.sy + L::cx(q))));
                else f[q] = LBM_LD(p.srcq[q] + ((mine && hint == 2) ? i + L::cx(q) * g.xl : i));
            }
        });
    } else {
        static_for<Q>([&](auto I) {
            constexpr int q = decltype(I)::value;
            f[q] = LBM_LD(p.srcq[q] + i);
        });
    }
    if constexpr (SPECULATIVE) {
        // The pulls must be REQUESTED before anything waits for the maps.  ptxas otherwise moves the "not streamed"
        // exit (which only needs the eagerly requested mask) in front of the pulls -- two dependent memory round
        // trips per cell instead of one.  An exit cannot move across a warp barrier.
        __syncwarp();
#if defined(LBM_HACK_NOWALL) && LBM_HACK_NOWALL == 2     /* timing experiment only (wrong results): no wall path at all */
        m = 0u;
#elif defined(LBM_HACK_NOWALL)                           /* ... no wall path for the cells of the two x faces */
        m = (((word >> (i & 31)) & 1u) && x != 1 && x != g.xl) ? p.mask[i] : 0u;
        if (m & MASK_SKIP) return;
#else
        m = ((word >> (i & 31)) & 1u) ? (face ? m_eager : p.mask[i]) : 0u;
        if (m & MASK_SKIP) return;
#endif
        if constexpr (MODE == SWEEP_XFACE) {
            if (hint != 0) {
                constexpr uint32_t QBITS = (1u << Q) - 1u;
                const uint32_t leaving = xs > 0 ? x_leaving_mask<Q>(1) : x_leaving_mask<Q>(-1);
                const uint32_t all_of_kind = hint == 1 ? MASK_ALLNOSLIP : MASK_ALLPERIODIC;
                if ((m & all_of_kind) && (m & QBITS) == leaving) {
                    m = 0;                                   // guessed right: f[] already holds the wall values
                } else {                                     // guessed wrong: pull the true sources, then the wall path as usual
                    static_for<Q>([&](auto I) {
                        constexpr int q = decltype(I)::value;
                        if constexpr (L::cx(q) != 0) {
                            if (xs == L::cx(q)) f[q] = LBM_LD(p.srcq[q] + i);
                        }
                    });
                }
            }
        }
    }
    finish_cell<Q, EXACT, MODE == SWEEP_SPLIT>(p, f, m, i, x, y, z);
}

// ---------------------------------------------------------------------------
// K1t: the same sweep fed by the TMA engine -- an OPT-IN engine (lbm_b200_set_sweep_engine / LBM_B200_TMA=1),
// bit-identical to sweep_kernel but slower: the load side alone streams 6.7 TB/s, but 16-24 consumer warps do
// not collide and store as many cells per second as sweep_kernel's 32 resident warps (profiles/variants_r07_tma.txt).  sweep_kernel's bytes in flight are bought with resident
// threads and land in registers (768-1024 threads x Q loads); a block spends part of its life computing,
// storing and being replaced, during which its registers hold nothing in flight.  Here ONE persistent
// block per SM keeps a ring of shared-memory stages (all ~220 KB of it) permanently in flight:
//   producer warp   per tile and population one cp.async.bulk.tensor of a BX x BY box of the 4-D tensor
//                   (x, y, z, q) over the source lattice; the pull offset is just the box origin
//                   (x0 - c_x, y0 - c_y, z - c_z, q) -- the copy engine does the shifted, unaligned read
//   consumer warps  wait for the stage, take their cell's Q values into registers, release the stage at
//                   once (before colliding), then run the SAME finish_cell as sweep_kernel: link-wise
//                   boundary values, BGK, Q coalesced 128-byte-aligned stores
// Tiles are dealt round-robin (tile t -> block t mod gridDim), so at any moment the chip works on one
// contiguous window of the lattice, like the hardware block scheduler does for sweep_kernel.
// The engine only accepts boxes that start on a 16-byte boundary (measured: an odd fp64 start coordinate is an
// illegal instruction, tools/selftest/tma_selftest.cu), so populations with c_x != 0 are fetched as a box that is
// two elements wider and starts one element early; the consumer reads at offset +1.  No extra 32-byte sector
// is touched by that: the wider box covers exactly the sectors the shifted row segment lives in.
// Out-of-range box elements (x beyond the row pitch) are zero-filled by the engine; they are only ever
// pull sources of flagged directions, which finish_cell replaces (never used while p.first is set).
template <int Q> struct TmaCfg {
    using L = Lattice<Q>;
    static constexpr int CELLS = 256;                         // cells per tile = consumer threads per group
    static constexpr int SLOT0 = CELLS * 8;                   // bytes of one population of a tile, c_x == 0
    static constexpr int SLOT1 = SLOT0 + 128;                 // c_x != 0: (BX+2) x BY doubles, rounded up to 128 bytes
    static constexpr int n_shifted()
    {
        int n = 0;
        for (int q = 0; q < Q; ++q) n += L::cx(q) != 0;
        return n;
    }
    static constexpr int slot_offset(int q)                   // bytes from the start of a stage
    {
        int o = 0;
        for (int k = 0; k < q; ++k) o += L::cx(k) != 0 ? SLOT1 : SLOT0;
        return o;
    }
    static constexpr int STAGE_BYTES = slot_offset(Q);
    static constexpr int tx_bytes(int bx) { return (Q - n_shifted()) * SLOT0 + n_shifted() * (bx + 2) * (CELLS / bx) * 8; }
#ifdef LBM_TMA_GROUPS
    static constexpr int GROUPS = LBM_TMA_GROUPS;
#else
    static constexpr int GROUPS = 2;                          // consumer groups working on alternate tiles
#endif
    // every group owns its own ring of DEPTH stages (a barrier is only ever waited on by ONE group, which
    // therefore can never be more than one phase ahead of it -- parity waits cannot alias)
#ifdef LBM_TMA_DEPTH
    static constexpr int DEPTH = LBM_TMA_DEPTH;
#else
    static constexpr int DEPTH = (227 * 1024 - 256) / STAGE_BYTES / GROUPS;
#endif
    static constexpr int STAGES = GROUPS * DEPTH;
    static constexpr int WARPS_PER_GROUP = CELLS / 32;
    static constexpr int CONSUMER_WARPS = GROUPS * WARPS_PER_GROUP;
    // A cp.async.bulk.tensor costs the ISSUING WARP ~150 ns, whatever its size (tools/selftest/tma_stream.cu:
    // one warp issuing 2 KB boxes feeds 2.0 TB/s chip-wide, 8 warps 6.0 TB/s, 19 warps 7.4 TB/s) -- so the
    // populations of a tile are fetched by several producer warps in parallel, population q by warp q mod N.
#ifdef LBM_TMA_PRODUCERS
    static constexpr int PRODUCER_WARPS = LBM_TMA_PRODUCERS;
#else
    static constexpr int PRODUCER_WARPS = 8;
#endif
    static constexpr int THREADS = (CONSUMER_WARPS + PRODUCER_WARPS) * 32;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 2 * STAGES * 8;
    static_assert(SMEM_BYTES <= 227 * 1024, "stage ring exceeds the shared memory of an SM");
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint64_t* bar)
{
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar)) : "memory");
}

template <int Q, bool EXACT, int BX>
__global__ void __launch_bounds__(TmaCfg<Q>::THREADS, 1)
sweep_tma_kernel(const SweepParams p, const __grid_constant__ CUtensorMap tmap0, const __grid_constant__ CUtensorMap tmap1,
                 const int tiles_x, const int tiles_y, const int nz)
{
    using C = TmaCfg<Q>;
    constexpr int BY = C::CELLS / BX;
    constexpr int S = C::STAGES;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t) S * C::STAGE_BYTES);
    uint64_t* empty = full + S;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const Layout& g = p.g;
    const unsigned int n_tiles = (unsigned int) tiles_x * tiles_y * nz;      // < 2^31 (checked by the host)
    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(full + s, C::PRODUCER_WARPS < Q ? C::PRODUCER_WARPS : Q);
            mbar_init(empty + s, C::WARPS_PER_GROUP);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp >= C::CONSUMER_WARPS) {
        // ---- producers: one thread per producer warp feeds populations pw, pw + N, pw + 2N, ...
        // (tmap0: BX x BY boxes, tmap1: (BX+2) x BY boxes for c_x != 0)
        const int pw = warp - C::CONSUMER_WARPS;
        if (lane != 0 || pw >= Q) return;
        const Tables<Q>& T = tables<Q>();
        // this warp's populations, their slots and box origins relative to the tile: fixed for the whole launch
        constexpr int MAX_OPS = (Q + C::PRODUCER_WARPS - 1) / C::PRODUCER_WARPS;
        int o_slot[MAX_OPS], o_x[MAX_OPS], o_y[MAX_OPS], o_z[MAX_OPS];
        const CUtensorMap* o_map[MAX_OPS];
        int my_bytes = 0;
        #pragma unroll
        for (int j = 0; j < MAX_OPS; ++j) {
            const int q = pw + j * C::PRODUCER_WARPS;
            o_slot[j] = 0; o_x[j] = o_y[j] = o_z[j] = 0; o_map[j] = &tmap0;
            if (q < Q) {
                for (int k = 0; k < q; ++k) o_slot[j] += T.c[k][0] != 0 ? C::SLOT1 : C::SLOT0;
                const int cxq = T.c[q][0];
                // box origin: x = 1 + tx*BX - c_x, one earlier if shifted; the tensor view starts TMA_X0 elements
                // into a row, where x = -1 would be -> the origin is always an even element (16-byte aligned)
                o_x[j] = 1 - cxq - (cxq != 0) + 1;
                o_y[j] = 1 - T.c[q][1];
                o_z[j] = p.z0 - T.c[q][2];
#ifdef LBM_TMA_ONEMAP      /* timing experiment only: wrong results */
                o_map[j] = &tmap0; o_x[j] = 2;
                my_bytes += C::SLOT0;
#else
                o_map[j] = cxq != 0 ? &tmap1 : &tmap0;
                my_bytes += cxq != 0 ? (BX + 2) * BY * 8 : C::SLOT0;
#endif
            }
        }
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap0) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap1) : "memory");
        int it = 0;
        for (unsigned int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
            const int k = it / C::GROUPS;                       // tile number inside its consumer group
            const int s = (it % C::GROUPS) * C::DEPTH + k % C::DEPTH;
            const uint32_t ph = (uint32_t) (k / C::DEPTH) & 1u;
            const unsigned int r = t / (unsigned int) tiles_x;
            const int tx = (int) (t - r * tiles_x);
            const int tz = (int) (r / (unsigned int) tiles_y), ty = (int) (r - tz * tiles_y);
            mbar_wait(empty + s, ph ^ 1u);           // a fresh barrier passes the wait on the opposite parity
            mbar_expect_tx(full + s, my_bytes);
            unsigned char* stage = smem_raw + (size_t) s * C::STAGE_BYTES;
            #pragma unroll
            for (int j = 0; j < MAX_OPS; ++j) {
                const int q = pw + j * C::PRODUCER_WARPS;
                if (q < Q) tma_load_4d(stage + o_slot[j], o_map[j], tx * BX + o_x[j], ty * BY + o_y[j], tz + o_z[j], q, full + s);
            }
        }
        return;
    }

    // ---- consumers: group `grp` takes every GROUPS-th tile of this block
    const int grp = warp / C::WARPS_PER_GROUP;
    const int cell = (warp % C::WARPS_PER_GROUP) * 32 + lane;      // position inside the tile, x fastest
    const int lx = cell % BX, ly = cell / BX;
    int it = 0;
    for (unsigned int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
        if (it % C::GROUPS != grp) continue;
        const int k = it / C::GROUPS;
        const int s = grp * C::DEPTH + k % C::DEPTH;
        const uint32_t ph = (uint32_t) (k / C::DEPTH) & 1u;
        const unsigned int r = t / (unsigned int) tiles_x;
        const int tx = (int) (t - r * tiles_x);
        const int tz = (int) (r / (unsigned int) tiles_y), ty = (int) (r - tz * tiles_y);
        const int x = 1 + tx * BX + lx, y = 1 + ty * BY + ly, z = p.z0 + tz;
        const bool inside = x <= g.xl && y <= g.yl;
        const int i = cell_at(g, x, y, z);
        const uint32_t word = inside ? p.bits[i >> 5] : 0u;        // requested before the wait: off the critical path
        mbar_wait(full + s, ph);
        const unsigned char* st = smem_raw + (size_t) s * C::STAGE_BYTES;
        double f[Q];
        static_for<Q>([&](auto I) {
            constexpr int q = decltype(I)::value;
            constexpr int off = C::slot_offset(q);
            if constexpr (Lattice<Q>::cx(q) != 0) f[q] = reinterpret_cast<const double*>(st + off)[ly * (BX + 2) + lx + 1];
            else f[q] = reinterpret_cast<const double*>(st + off)[cell];
        });
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + s);                     // the stage can be refilled while we collide
        if (!inside) continue;
        const uint32_t m = ((word >> (i & 31)) & 1u) ? p.mask[i] : 0u;
        if (m & MASK_SKIP) continue;
#ifdef LBM_TMA_NOSTORE      /* timing experiment only: wrong results */
        double acc = 0.0;
        static_for<Q>([&](auto I) { acc += f[decltype(I)::value]; });
        if (acc == 1.2345) p.dstq[0][i] = acc;
#else
        finish_cell<Q, EXACT>(p, f, m, i, x, y, z);
#endif
    }
}

// K1g: ghost-shell cells that kept the fluid handler are BGK-collided in place in
// the field that has just become the collide field (domain.hpp:147-155); they are
// never streamed into (domain.hpp:121-123).
template <int Q, bool EXACT>
__global__ void ghost_fluid_kernel(double* __restrict__ field, long long qstride, const int* __restrict__ cells,
                                   int n, double tau, double omega)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int i = cells[t];
    double f[Q];
    static_for<Q>([&](auto I) { constexpr int q = decltype(I)::value; f[q] = field[q * qstride + i]; });
    bgk_collide<Q, EXACT>(f, tau, omega);
    static_for<Q>([&](auto I) { constexpr int q = decltype(I)::value; field[q * qstride + i] = f[q]; });
}

// K2: the reference's non-fluid pass (domain.hpp:157-165) on the current collide
// field, so that boundary cells read back through Domain::cell()/VTK hold what
// the reference holds.  Reads fluid cells, writes non-fluid cells: no hazard.
template <int Q, bool EXACT>
__global__ void materialize_kernel(double* __restrict__ field, const uint8_t* __restrict__ kind,
                                   const uint16_t* __restrict__ bcid, const BcRec* __restrict__ bc, const Layout g,
                                   const int lo_interface, const int hi_interface, const int z_lo_open, const int z_hi_open)
{
    // lo/hi_interface: the ghost plane on that side is a slab interface, its cells are interior cells of the
    // neighbour slab.  z_lo_open / z_hi_open: the neighbour's FULL edge plane was pushed there for this time
    // level (lbm_b200_halo_push_all), so links from our boundary cells into it can be evaluated too.
    // Boundary cells INSIDE an interface ghost plane are the neighbour's, but their entries that point into
    // our own interior are read by our edge-plane cells when the next sweep pulls stored values (`first`);
    // those depend on our own cells only and are always written here.
    using L = Lattice<Q>;
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    const int z = blockIdx.z;
    if (x > g.xl + 1) return;
    const int slow = g.swap ? y : z, ns = n_slow(g), nm = n_mid(g);
    const bool in_interface = (slow == 0 && lo_interface) || (slow == ns + 1 && hi_interface);
    const int b = cell_at(g, x, y, z);
    const int k = kind[b];
    if (k < K_NOSLIP || k > K_PRESSURE) return;
    const BcRec* rec = bc + bcid[b];
    static_for<Q>([&](auto I) {
        constexpr int q = decltype(I)::value;
        const int nx = x + L::cx(q), ny = y + L::cy(q), nz = z + L::cz(q);
        const int n_s = g.swap ? ny : nz, n_m = g.swap ? nz : ny;       // neighbour along the slow / middle axis
        bool inb = nx > 0 && nx < g.xl + 1 && n_m > 0 && n_m < nm + 1;
        if (in_interface) inb = inb && n_s > 0 && n_s < ns + 1;
        else inb = inb && (n_s > 0 || (n_s == 0 && z_lo_open)) && (n_s < ns + 1 || (n_s == ns + 1 && z_hi_open));
        if (inb) {
            const int n = b + (L::cz(q) * g.sz + L::cy(q) * g.sy + L::cx(q));
            if (kind[n] == K_FLUID) {
                OwnMoments om;
                om.have = false;
                om.rec_of_u = nullptr;
        om.rec_of_u = nullptr;
                field[q * g.qstride + b] = link_value<Q, EXACT, q>(field, kind, g, n, k, rec, om);
            }
        }
    });
}

// K3: io/vtk.hpp:62-73 -> compute_density / compute_velocity (collision.hpp:7-31),
// always in the reference's association.  Dense outputs in z,y,x order.
template <int Q>
__global__ void macroscopic_kernel(const double* __restrict__ field, const Layout g,
                                   double* __restrict__ rho_out, double* __restrict__ u_out, const int z_begin)
{
    const int x = 1 + blockIdx.x * blockDim.x + threadIdx.x;
    const int y = 1 + blockIdx.y;
    const int z = 1 + z_begin + blockIdx.z;
    if (x > g.xl) return;
    const int i = cell_at(g, x, y, z);
    double f[Q];
    static_for<Q>([&](auto I) { constexpr int q = decltype(I)::value; f[q] = field[q * g.qstride + i]; });
    double rho, mx, my, mz;
    moments_exact<Q>(f, rho, mx, my, mz);
    const long long o = ((long long) (z - 1) * g.yl + (y - 1)) * g.xl + (x - 1);
    if (rho_out) rho_out[o] = rho;
    if (u_out) {
        u_out[3 * o + 0] = __ddiv_rn(mx, rho);
        u_out[3 * o + 1] = __ddiv_rn(my, rho);
        u_out[3 * o + 2] = __ddiv_rn(mz, rho);
    }
}

// per-block partial sums over interior fluid cells: mass, sum |u|^2, max |u|^2
template <int Q>
__global__ void diagnostics_kernel(const double* __restrict__ field, const uint8_t* __restrict__ kind, const Layout g,
                                   double* __restrict__ partial /* [nblocks][3] */)
{
    __shared__ double sh[3][128];
    const int x = 1 + blockIdx.x * blockDim.x + threadIdx.x;
    const int y = 1 + blockIdx.y;
    const int z = 1 + blockIdx.z;
    double mass = 0.0, ke = 0.0, um = 0.0;
    if (x <= g.xl) {
        const int i = cell_at(g, x, y, z);
        if (kind[i] == K_FLUID) {
            double f[Q];
            static_for<Q>([&](auto I) { constexpr int q = decltype(I)::value; f[q] = field[q * g.qstride + i]; });
            double rho, mx, my, mz;
            moments_exact<Q>(f, rho, mx, my, mz);
            const double ux = mx / rho, uy = my / rho, uz = mz / rho;
            mass = rho;
            ke = ux * ux + uy * uy + uz * uz;
            um = ke;
        }
    }
    sh[0][threadIdx.x] = mass; sh[1][threadIdx.x] = ke; sh[2][threadIdx.x] = um;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if ((int) threadIdx.x < s) {
            sh[0][threadIdx.x] += sh[0][threadIdx.x + s];
            sh[1][threadIdx.x] += sh[1][threadIdx.x + s];
            sh[2][threadIdx.x] = fmax(sh[2][threadIdx.x], sh[2][threadIdx.x + s]);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const long long blk = ((long long) blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
        partial[3 * blk + 0] = sh[0][0];
        partial[3 * blk + 1] = sh[1][0];
        partial[3 * blk + 2] = sh[2][0];
    }
}

// ---------------------------------------------------------------------------
// K5: set-up kernels

// Cell ctor, cell.hpp:9-15: pdf = weights (padding included, harmless)
template <int Q>
__global__ void fill_weights_kernel(double* __restrict__ field, long long qstride)
{
    const Tables<Q>& T = tables<Q>();
    const long long n = qstride * Q;
    for (long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x; i < n; i += (long long) gridDim.x * blockDim.x)
        field[i] = T.w[i / qstride];
}

// Link mask of the step  src lattice -> dst lattice.  Handlers belong to lattices (cell.h:15), and the
// reference decides "is streamed" with the source lattice's handlers (domain.hpp:124, before the swap) and
// "is collided" with the destination's (domain.hpp:148-165, after it).  Normally both lattices carry the
// same handlers (kind_src == kind_dst).
//   bit q           the pull source X - c_q of streamed cell X is not fluid (in the source lattice)
//   MASK_SKIP       X is not streamed
//   MASK_NOCOLLIDE  X is streamed but not BGK-collided
// counters[0] += cells that are collided in place without being streamed (fluid in the destination lattice,
// but in the ghost shell or not fluid in the source lattice; domain.hpp:147-155 loops 0..l+1);
// counters[1] |= 1 if a ghost plane of the slow axis that belongs to the physical shell carries PERIODIC;
// counters[3] += interior cells that are not streamed (solid in the source lattice).
template <int Q>
__global__ void build_mask_kernel(const uint8_t* __restrict__ kind_src, const uint8_t* __restrict__ kind_dst,
                                  const uint16_t* __restrict__ bcid_src,
                                  uint32_t* __restrict__ mask, uint32_t* __restrict__ bits, const Layout g,
                                  const int lo_interface, const int hi_interface, unsigned int* __restrict__ counters)
{
    const Tables<Q>& T = tables<Q>();
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    const int z = blockIdx.z;
    if (x > g.xl + 1) return;
    const int i = cell_at(g, x, y, z);
    const bool interior = x > 0 && x < g.xl + 1 && y > 0 && y < g.yl + 1 && z > 0 && z < g.zl + 1;
    const bool streamed = interior && kind_src[i] == K_FLUID;
    uint32_t m = 0;
    if (!streamed) {
        m = MASK_SKIP;
    } else {
        bool only_noslip = true, only_periodic = true, one_handler = true;
        int k0 = -1, id0 = -1;
        for (int q = 0; q < Q; ++q) {
            const int s = i - (T.c[q][2] * g.sz + T.c[q][1] * g.sy + T.c[q][0]);
            const int k = kind_src[s];
            if (k != K_FLUID) {
                m |= 1u << q;
                only_noslip = only_noslip && k == K_NOSLIP;
                only_periodic = only_periodic && k == K_PERIODIC;
                const int id = bcid_src[s];
                if (k0 < 0) { k0 = k; id0 = id; }
                one_handler = one_handler && k == k0 && id == id0;
            }
        }
        // (FreeSlip looks at the neighbours of each boundary cell, Null / Parallel / Periodic use stored values)
        if (m && one_handler && (k0 == K_MOVINGWALL || k0 == K_INFLOW || k0 == K_OUTFLOW || k0 == K_PRESSURE)) m |= MASK_ONEHANDLER;
        if (m && only_noslip) m |= MASK_ALLNOSLIP;
        if (m && only_periodic) m |= MASK_ALLPERIODIC;
        if (kind_dst[i] != K_FLUID) m |= MASK_NOCOLLIDE;
    }
    mask[i] = m;
    if (m) atomicOr(&bits[i >> 5], 1u << (i & 31));   // bits zeroed by the caller
    // x-face statistics for SweepParams::xhint_*: cells (not on a y / z face) of the x = 1 / x = xl face whose flagged
    // links are exactly the c_x = +1 / -1 ones and all NoSlipBoundary (counters[8], [10]) or all PERIODIC ([9], [11])
    if (streamed && (x == 1 || x == g.xl) && y > 1 && y < g.yl && z > 1 && z < g.zl) {
        const int sign = x == 1 ? 1 : -1;
        uint32_t leaving = 0;
        for (int q = 0; q < Q; ++q)
            if (T.c[q][0] == sign) leaving |= 1u << q;
        if ((m & ((1u << Q) - 1u)) == leaving) {
            if (m & MASK_ALLNOSLIP) atomicAdd(&counters[x == 1 ? 8 : 10], 1u);
            if (m & MASK_ALLPERIODIC) atomicAdd(&counters[x == 1 ? 9 : 11], 1u);
        }
    }
    if ((m & MASK_SKIP) && interior) atomicAdd(&counters[3], 1u);
    const int slow = g.swap ? y : z;
    const bool neighbours_cell = (slow == 0 && lo_interface) || (slow == n_slow(g) + 1 && hi_interface);
    if (!neighbours_cell) {
        if (!streamed && kind_dst[i] == K_FLUID) atomicAdd(&counters[0], 1u);
        if ((slow == 0 || slow == n_slow(g) + 1) && kind_src[i] == K_PERIODIC) atomicOr(&counters[1], 1u);
    }
}

// second pass, only when the first one counted something: the list of in-place collided cells
__global__ void collect_inplace_kernel(const uint8_t* __restrict__ kind_src, const uint8_t* __restrict__ kind_dst,
                                       const Layout g, const int lo_interface, const int hi_interface,
                                       int* __restrict__ list, unsigned int capacity, unsigned int* __restrict__ cursor)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    const int z = blockIdx.z;
    if (x > g.xl + 1) return;
    const int slow = g.swap ? y : z;
    if ((slow == 0 && lo_interface) || (slow == n_slow(g) + 1 && hi_interface)) return;
    const int i = cell_at(g, x, y, z);
    const bool interior = x > 0 && x < g.xl + 1 && y > 0 && y < g.yl + 1 && z > 0 && z < g.zl + 1;
    const bool streamed = interior && kind_src[i] == K_FLUID;
    if (!streamed && kind_dst[i] == K_FLUID) {
        const unsigned int slot = atomicAdd(cursor, 1u);
        if (slot < capacity) list[slot] = i;
    }
}

// Domain::setBoundaryCondition (domain.hpp:185-193) for one inclusive box in LOCAL coordinates
__global__ void paint_box_kernel(uint8_t* __restrict__ kind, uint16_t* __restrict__ bcid, const Layout g,
                                 int x0, int y0, int z0, int nx, int ny, uint8_t k, uint16_t id)
{
    const int dx = blockIdx.x * blockDim.x + threadIdx.x;
    if (dx >= nx) return;
    const int i = cell_at(g, x0 + dx, y0 + blockIdx.y, z0 + blockIdx.z);
    kind[i] = k;
    bcid[i] = id;
}

// the loop of io/vtk.hpp:141-150: mask value 0 => the solid handler.  `mask` holds x-y planes of interior cells,
// `mask_rows` rows each; local cell (x, y, z) takes mask element (x-1, y - y_shift, z - z_shift).  Painted: local
// rows [y_lo, y_lo + gridDim.y) of local planes [z_lo, z_lo + gridDim.z).
__global__ void paint_mask_kernel(const uint8_t* __restrict__ mask, uint8_t* __restrict__ kind, uint16_t* __restrict__ bcid,
                                  const Layout g, int y_lo, int z_lo, int mask_rows, int y_shift, int z_shift, uint8_t k, uint16_t id)
{
    const int x = 1 + blockIdx.x * blockDim.x + threadIdx.x;
    const int y = y_lo + blockIdx.y;
    const int z = z_lo + blockIdx.z;
    if (x > g.xl) return;
    const size_t o = ((size_t) (z - z_shift) * mask_rows + (y - y_shift)) * g.xl + (x - 1);
    if (!mask[o]) {
        const int i = cell_at(g, x, y, z);
        kind[i] = k;
        bcid[i] = id;
    }
}

// Domain::set_nonfluid_cells_nullcollide (domain.hpp:101-113, cell.hpp:75-92): interior cells without an
// in-bounds fluid neighbour (the q loop includes the rest velocity, so fluid cells never qualify) take the
// do-nothing handler.  Tagging never changes who is fluid, so it can be done in place.
template <int Q>
__global__ void tag_null_kernel(uint8_t* __restrict__ kind, const Layout g, int slow_first, int slow_global,
                                unsigned int* __restrict__ count)
{
    const Tables<Q>& T = tables<Q>();
    const int x = 1 + blockIdx.x * blockDim.x + threadIdx.x;
    const int y = 1 + blockIdx.y;
    const int z = 1 + blockIdx.z;
    if (x > g.xl) return;
    const int i = cell_at(g, x, y, z);
    if (kind[i] == K_FLUID) return;
    for (int q = 0; q < Q; ++q) {
        const int nx = x + T.c[q][0], ny = y + T.c[q][1], nz = z + T.c[q][2];
        const int n_m = g.swap ? nz : ny;                                  // middle axis: local == global
        const int gs = (g.swap ? ny : nz) + slow_first - 1;               // slow axis: global index of the neighbour
        if (nx > 0 && nx < g.xl + 1 && n_m > 0 && n_m < n_mid(g) + 1 && gs > 0 && gs < slow_global + 1
            && kind[i + (T.c[q][2] * g.sz + T.c[q][1] * g.sy + T.c[q][0])] == K_FLUID)
            return;
    }
    if (kind[i] != K_NULL) atomicAdd(count, 1u);
    kind[i] = K_NULL;
}

// a tagged cell keeps its former handler id: restore its kind from the table (used when the two lattices
// start to differ and only one of them carries the tags)
__global__ void untag_null_kernel(uint8_t* __restrict__ kind, const uint16_t* __restrict__ bcid,
                                  const BcRec* __restrict__ bc, int n_bc, long long n)
{
    const long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (kind[i] == K_NULL && bcid[i] < n_bc) kind[i] = (uint8_t) bc[bcid[i]].kind;
}

// Check of dense maps (reference idx order, planes [z0, z0+nz) of the slab) before they are accepted:
// known kinds, handler ids inside the table and of the same kind, PERIODIC only on the ghost shell;
// result[0] = min over offending cells of (dense index * 8 + reason)
__global__ void validate_dense_kernel(const uint8_t* __restrict__ kind, const uint16_t* __restrict__ bcid,
                                      const BcRec* __restrict__ bc, int n_bc, const Layout g, int z0,
                                      unsigned long long* __restrict__ result)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    const int zz = blockIdx.z;
    if (x > g.xl + 1) return;
    const unsigned long long d = ((unsigned long long) zz * (g.yl + 2) + y) * (g.xl + 2) + x;
    const int k = kind[d];
    const int z = z0 + zz;
    int why = 0;
    if (k >= K_COUNT) why = 1;
    else if (k >= K_NOSLIP && k <= K_PRESSURE) {
        const int id = bcid[d];
        if (id >= n_bc) why = 2;
        else if (bc[id].kind != k) why = 3;
    } else if (k == K_PERIODIC && x > 0 && x < g.xl + 1 && y > 0 && y < g.yl + 1 && z > 0 && z < g.zl + 1) why = 4;
    if (why) atomicMin(result, (d + (unsigned long long) z0 * (g.yl + 2) * (g.xl + 2)) * 8ull + (unsigned long long) why);
}

// padded device maps -> dense (reference idx order), planes [z0, z0+nz).  report_former: cells tagged by
// set_nonfluid_cells_nullcollide report their former handler's kind (see lbm_b200_tag_null_cells).
__global__ void gather_maps_kernel(const uint8_t* __restrict__ kind, const uint16_t* __restrict__ bcid,
                                   const BcRec* __restrict__ bc, int n_bc, uint8_t* __restrict__ kind_out,
                                   uint16_t* __restrict__ bcid_out, const Layout g, int z0, int report_former)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    const int zz = blockIdx.z;
    if (x > g.xl + 1) return;
    const int i = cell_at(g, x, y, z0 + zz);
    const long long d = ((long long) zz * (g.yl + 2) + y) * (g.xl + 2) + x;
    int k = kind[i];
    const int id = bcid[i];
    if (report_former && k == K_NULL && id < n_bc) k = bc[id].kind;
    if (kind_out) kind_out[d] = (uint8_t) k;
    if (bcid_out) bcid_out[d] = (uint16_t) id;
}

// dense (reference idx order) byte/short maps -> padded device maps, planes [z0, z0+nz)
template <typename T>
__global__ void scatter_map_kernel(const T* __restrict__ dense, T* __restrict__ padded, const Layout g, int z0)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    const int zz = blockIdx.z;
    if (x > g.xl + 1) return;
    padded[cell_at(g, x, y, z0 + zz)] = dense[((long long) zz * (g.yl + 2) + y) * (g.xl + 2) + x];
}

// AoS chunk (reference Cell order, planes [z0, z0+nz)) <-> padded SoA
template <int Q, bool TO_DEVICE_LAYOUT>
__global__ void transpose_aos_kernel(double* __restrict__ aos, double* __restrict__ field, const Layout g, int z0)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    const int zz = blockIdx.z;
    if (x > g.xl + 1) return;
    const int i = cell_at(g, x, y, z0 + zz);
    double* a = aos + (((long long) zz * (g.yl + 2) + y) * (g.xl + 2) + x) * Q;
    #pragma unroll
    for (int q = 0; q < Q; ++q) {
        if (TO_DEVICE_LAYOUT) field[q * g.qstride + i] = a[q];
        else a[q] = field[q * g.qstride + i];
    }
}

// f = feq(rho,u) with compute_feq's association (collision.hpp:34-51), planes [z0,z0+nz)
template <int Q>
__global__ void equilibrium_kernel(const double* __restrict__ rho, const double* __restrict__ u,
                                   double* __restrict__ field, const Layout g, int z0)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    const int zz = blockIdx.z;
    if (x > g.xl + 1) return;
    const int i = cell_at(g, x, y, z0 + zz);
    const long long d = ((long long) zz * (g.yl + 2) + y) * (g.xl + 2) + x;
    const double r = rho[d], ux = u[3 * d], uy = u[3 * d + 1], uz = u[3 * d + 2];
    const double uut = uu_term_exact(ux, uy, uz);
    static_for<Q>([&](auto I) {
        constexpr int q = decltype(I)::value;
        field[q * g.qstride + i] = feq_exact<Q, q>(r, ux, uy, uz, uut);
    });
}

// Self-test of div_rcp against IEEE division (lbm_b200_selftest_division): operand n of divisor set `which`
// (0: C_S^2, 1: 2 C_S^4, 2: 2 C_S^2, 3: tau, 4: a random rho-like divisor per operand) is a hash of (seed, n) with a
// random 52-bit mantissa and an exponent in 2^-60 .. 2^4 -- the magnitudes the sweep divides.
__device__ __forceinline__ uint64_t mix64(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__global__ void division_selftest_kernel(unsigned long long n, unsigned long long seed, int which, double tau, double rcp_tau,
                                         unsigned long long* __restrict__ mismatches)
{
    unsigned long long bad = 0;
    for (unsigned long long k = blockIdx.x * (unsigned long long) blockDim.x + threadIdx.x; k < n;
         k += (unsigned long long) gridDim.x * blockDim.x) {
        const uint64_t h = mix64(seed * 0x100000001B3ull + k), h2 = mix64(h);
        const uint64_t e = 1023 - 60 + (h2 >> 8) % 65;
        const double a = __longlong_as_double((long long) (((h2 & 1) << 63) | (e << 52) | (h >> 12)));
        double b, y;
        switch (which) {
        case 0: b = CS2; y = RCP_CS2; break;
        case 1: b = TWO_CS4; y = RCP_TWO_CS4; break;
        case 2: b = TWO_CS2; y = RCP_TWO_CS2; break;
        case 3: b = tau; y = rcp_tau; break;
        default: {
            const uint64_t h3 = mix64(h2);
            b = __longlong_as_double((long long) (((uint64_t) (1023 - 2 + (h3 >> 60) % 4) << 52) | (h3 >> 12)));
            y = __drcp_rn(b);
        }
        }
        if (__double_as_longlong(div_rcp(a, b, y)) != __double_as_longlong(__ddiv_rn(a, b))) ++bad;
    }
    if (bad) atomicAdd(mismatches, bad);
}

// ---------------------------------------------------------------------------
// K4: slab hand-shake for direct peer stores.  Each slab owns two 64-bit arrival
// counters (one per side) that the neighbour on that side bumps, with system
// scope, after its sweep has stored its leaving populations into our ghost plane.
// Before sweep number n+1 a slab waits until both neighbours have completed n
// sweeps: that orders "their halo stores landed" (read-after-write) as well as
// "they no longer read the buffer we are about to overwrite" (write-after-read).
// The epoch (sweeps this slab has completed since its peers were connected) lives in device memory and is
// advanced by the signal kernel itself, so the wait / sweep / signal sequence of a step has no host-supplied
// argument that changes from step to step and can be replayed from a CUDA graph.
__global__ void halo_signal_kernel(unsigned long long* peer_flag_a, unsigned long long* peer_flag_b,
                                   unsigned long long* my_epoch, unsigned long long* scratch)
{
    // atomics are performed at the owning GPU's L2 (the point of coherence), so the value is visible
    // to the neighbour's poll as soon as the NVLink transaction lands.  The RETURNING form is used on
    // purpose: it is a round trip, so this kernel only retires once the neighbour really has the
    // value (a posted reduction may linger in the fabric until later traffic pushes it along).
    __threadfence_system();
    const unsigned long long epoch = *my_epoch + 1;
    *my_epoch = epoch;
    unsigned long long seen = 0;
    if (peer_flag_a) seen += atomicMax_system(peer_flag_a, epoch);
    if (peer_flag_b) seen += atomicMax_system(peer_flag_b, epoch);
    *scratch = seen;
}

__global__ void halo_wait_kernel(unsigned long long* flag_a, unsigned long long* flag_b,
                                 const unsigned long long* my_epoch, long long timeout_cycles, int* error_word,
                                 volatile int* host_error_word, unsigned long long* trace, int trace_epochs)
{
    const unsigned long long epoch = *my_epoch;
    // optional trace (LBM_B200_HALO_TRACE): nanosecond timestamps of entry and exit per epoch
    if (trace && epoch < (unsigned long long) trace_epochs) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(trace[2 * epoch]));
    const long long t0 = clock64();
    unsigned long long* flags[2] = { flag_a, flag_b };
    for (int k = 0; k < 2; ++k) {
        if (!flags[k]) continue;
        for (;;) {
            // read-modify-write read: served by L2, never by a stale cached copy
            const unsigned long long v = atomicAdd_system(flags[k], 0ull);
            if (v >= epoch) break;
            if (clock64() - t0 > timeout_cycles) {   // never hang the GPU on a lost neighbour
                *error_word = 1;
                if (host_error_word) *host_error_word = 1;   // pinned mirror: the host refuses further steps at once
                return;
            }
            __nanosleep(100);
        }
    }
    __threadfence_system();
    if (trace && epoch < (unsigned long long) trace_epochs) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(trace[2 * epoch + 1]));
}

} // namespace lbmb200

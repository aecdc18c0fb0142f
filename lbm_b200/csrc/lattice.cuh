// lattice.cuh -- compile-time lattice descriptors for device and host code.
//
// Same content as the reference's model.h (d3q15 :13-46, d3q19 :53-87, d3q27
// :96-134): velocity order, weights, inv(q) = Q-1-q.  The reference tables are
// the 27 vectors of {-1,0,1}^3 in z-slowest / x-fastest order restricted by
// |c|^2, so they are generated here by constexpr enumeration; q is always a
// compile-time constant in the kernels (static_for), which turns c_q and w_q
// into immediates -- the "lattice weights and velocities in constant memory" of
// the north star, one level better.
#pragma once
#include <utility>
#include <type_traits>

#if defined(__CUDACC__)
#define LBM_HD __host__ __device__
#else
#define LBM_HD
#endif

namespace lbmb200 {

// lbmdefinitions.h:47 -- the reference's truncated speed of sound.  C_S*C_S is
// 0.33333333333376547, NOT 1/3; every derived constant keeps the reference's
// left-to-right association (collision.hpp:47-48).
constexpr double C_S = 0.57735026919;
constexpr double CS2 = C_S * C_S;                      // (C_S * C_S)
constexpr double TWO_CS4 = 2 * C_S * C_S * C_S * C_S;  // (((2*C_S)*C_S)*C_S)*C_S
constexpr double TWO_CS2 = 2 * C_S * C_S;              // (2*C_S)*C_S

template <int Q>
struct Lattice {
    static_assert(Q == 15 || Q == 19 || Q == 27, "D3Q15, D3Q19 or D3Q27");

    LBM_HD static constexpr bool keep(int n2)
    {
        return Q == 27 ? true : (Q == 19 ? n2 <= 2 : n2 != 2);
    }
    // position (0..26) of the q-th kept vector inside the full cube enumeration
    LBM_HD static constexpr int cube(int q)
    {
        int k = 0;
        for (int i = 0; i < 27; ++i) {
            const int x = i % 3 - 1, y = (i / 3) % 3 - 1, z = i / 9 - 1;
            if (keep(x * x + y * y + z * z)) {
                if (k == q) return i;
                ++k;
            }
        }
        return -1;
    }
    LBM_HD static constexpr int cx(int q) { return cube(q) % 3 - 1; }
    LBM_HD static constexpr int cy(int q) { return (cube(q) / 3) % 3 - 1; }
    LBM_HD static constexpr int cz(int q) { return cube(q) / 9 - 1; }
    LBM_HD static constexpr int norm2(int q)
    {
        return cx(q) * cx(q) + cy(q) * cy(q) + cz(q) * cz(q);
    }
    LBM_HD static constexpr double w(int q)
    {
        const int n = norm2(q);
        if (Q == 15) return n == 0 ? 16.0 / 72 : (n == 1 ? 8.0 / 72 : 1.0 / 72);
        if (Q == 19) return n == 0 ? 12.0 / 36 : (n == 1 ? 2.0 / 36 : 1.0 / 36);
        return n == 0 ? 64.0 / 216 : (n == 1 ? 16.0 / 216 : (n == 2 ? 4.0 / 216 : 1.0 / 216));
    }
    LBM_HD static constexpr int inv(int q) { return Q - 1 - q; }
    // index of velocity (u,v,w), -1 if it is not in the set; agrees with the
    // reference's velocity_index formulas (model.h:43-46, 83-86, 130-133) on
    // every member of the set
    LBM_HD static constexpr int index_of(int u, int v, int ww)
    {
        for (int q = 0; q < Q; ++q)
            if (cx(q) == u && cy(q) == v && cz(q) == ww) return q;
        return -1;
    }
    // number of populations with c_z = +1 (== those with c_z = -1): 5/5/9
    LBM_HD static constexpr int n_up()
    {
        int n = 0;
        for (int q = 0; q < Q; ++q) n += cz(q) == 1;
        return n;
    }
};

template <typename F, int... I>
LBM_HD inline __attribute__((always_inline)) void static_for_impl(F&& f, std::integer_sequence<int, I...>)
{
    (f(std::integral_constant<int, I>{}), ...);
}
// calls f(integral_constant<int,0>) ... f(integral_constant<int,N-1>)
template <int N, typename F>
LBM_HD inline __attribute__((always_inline)) void static_for(F&& f)
{
    static_for_impl(static_cast<F&&>(f), std::make_integer_sequence<int, N>{});
}

} // namespace lbmb200

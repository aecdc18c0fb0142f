// include/lbm/model.h -- lattice descriptors d3q15 / d3q19 / d3q27.
//
// Public members are those of the reference's include/model.h (:13-46, :53-87,
// :96-134): D, Q, name, velocities[q][d] (double, like the reference), weights[q],
// inv(q), velocity_index(u,v,w).  The tables are not typed in: the reference's
// orderings are the cube {-1,0,1}^3 enumerated z-slowest / x-fastest and filtered
// by |c|^2, and the weights depend on |c|^2 only, so they are generated at compile
// time.  tests/test_host_surface.py compares every entry with the reference tables.
// The device code uses the same generator (lbm_b200/csrc/lattice.cuh).
#pragma once
#include <array>
#include <cstddef>
#include "lbmdefinitions.h"

namespace lbm
{
namespace model
{
namespace detail
{

template <std::size_t QN>
struct Tables {
    std::array<std::array<double, 3>, QN> velocities;
    std::array<double, QN> weights;
    std::array<int, 27> lookup;   // cube position -> q, or -1
};

// norm_weights[n] = weight of a velocity with |c|^2 == n, or 0 if such velocities are not in the set
template <std::size_t QN>
constexpr Tables<QN> generate(const std::array<double, 4>& norm_weights)
{
    Tables<QN> t{};
    std::size_t q = 0;
    for (int i = 0; i < 27; ++i) {
        const int x = i % 3 - 1, y = (i / 3) % 3 - 1, z = i / 9 - 1;
        const int n = x * x + y * y + z * z;
        t.lookup[i] = -1;
        if (norm_weights[n] == 0.0) continue;
        t.velocities[q] = { double(x), double(y), double(z) };
        t.weights[q] = norm_weights[n];
        t.lookup[i] = int(q);
        ++q;
    }
    return t;
}

template <typename Self, std::size_t QN>
struct Descriptor {
    static constexpr std::size_t D = 3;
    static constexpr std::size_t Q = QN;
    // inv(q): the tables are point-symmetric, so the opposite velocity is Q-1-q
    static int inv(int q) { return int(QN) - 1 - q; }
    static std::size_t velocity_index(int u, int v, int w)
    {
        return std::size_t(Self::tables.lookup[(w + 1) * 9 + (v + 1) * 3 + (u + 1)]);
    }
};

} // namespace detail

struct d3q15 : detail::Descriptor<d3q15, 15> {
    static constexpr const char* const name = "D3Q15";
    static constexpr detail::Tables<15> tables = detail::generate<15>({ 16.0 / 72, 8.0 / 72, 0.0, 1.0 / 72 });
    static constexpr const std::array<std::array<double, 3>, 15>& velocities = tables.velocities;
    static constexpr const std::array<double, 15>& weights = tables.weights;
};

struct d3q19 : detail::Descriptor<d3q19, 19> {
    static constexpr const char* const name = "D3Q19";
    static constexpr detail::Tables<19> tables = detail::generate<19>({ 12.0 / 36, 2.0 / 36, 1.0 / 36, 0.0 });
    static constexpr const std::array<std::array<double, 3>, 19>& velocities = tables.velocities;
    static constexpr const std::array<double, 19>& weights = tables.weights;
};

struct d3q27 : detail::Descriptor<d3q27, 27> {
    static constexpr const char* const name = "D3Q27";
    static constexpr detail::Tables<27> tables =
            detail::generate<27>({ 64.0 / 216, 16.0 / 216, 4.0 / 216, 1.0 / 216 });
    static constexpr const std::array<std::array<double, 3>, 27>& velocities = tables.velocities;
    static constexpr const std::array<double, 27>& weights = tables.weights;
};

} // namespace model
} // namespace lbm

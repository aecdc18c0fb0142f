// include/lbm/lbmdefinitions.h -- type aliases and C_S of the B200 host surface.
// Same names as the reference's include/lbmdefinitions.h:12-47 so that caller code
// (src/main.cpp, io/scenario.h style set-up code) compiles unchanged.
#pragma once
#include <array>
#include <cstddef>
#include <cstdint>
#include <memory>
#include <vector>

namespace lbm
{

template <std::size_t Q, std::size_t D> using lattice_velocities = std::array<std::array<int, D>, Q>;
template <std::size_t N> using lattice_weights = std::array<double, N>;
template <std::size_t N> using double_array = std::array<double, N>;
template <std::size_t N> using int_array = std::array<int, N>;
template <std::size_t N> using uint_array = std::array<std::uint64_t, N>;

template <typename lattice_model> class Cell;
// In the reference this is the storage of a lattice (vector of Cell structs,
// lbmdefinitions.h:28-29).  Here populations live on the GPU as a padded
// structure of arrays; a Lattice_field only exists as the lazily synchronised host
// mirror behind Domain::cell().
template <typename lattice_model> using Lattice_field = std::vector<Cell<lattice_model>>;

template <typename lattice_model> class Domain;
template <typename lattice_model> using Domain_ptr = std::unique_ptr<Domain<lattice_model>>;

template <typename lattice_model> class FluidCollision;
template <typename lattice_model> using FluidColl_ptr = std::shared_ptr<FluidCollision<lattice_model>>;
template <typename lattice_model> class NonFluidCollision;
template <typename lattice_model> using NonFluidColl_ptr = std::shared_ptr<NonFluidCollision<lattice_model>>;

// The reference's truncated lattice speed of sound (lbmdefinitions.h:47).  Kept to
// the digit: C_S*C_S = 0.33333333333376547 differs from 1/3 by 1.3e-12 relative,
// which is the whole parity budget.
static constexpr double C_S = 0.57735026919l;

} // namespace lbm

// include/lbm/cell.h -- Cell<M>: one lattice site of the HOST MIRROR.
//
// Interface of the reference's include/cell.h:12-69.  In the reference a Cell is
// the storage itself (Q doubles + handler pointer, cell.h:14-15); here the
// populations live on the GPU and a Cell is an element of the mirror that
// Domain::cell() synchronises lazily (download on first read after a step,
// upload before the next step if a non-const reference was handed out).
#pragma once
#include <algorithm>
#include <cstddef>

#include "collision.h"

namespace lbm
{

template <typename lattice_model>
class Cell
{
    double_array<lattice_model::Q> pdf;
    const Collision<lattice_model>* collision;

public:
    // initial populations are the lattice weights: density 1, velocity 0 (cell.hpp:9-15)
    explicit Cell(const Collision<lattice_model>* collision) : collision { collision }
    {
        for (std::size_t q = 0; q < lattice_model::Q; ++q) pdf[q] = lattice_model::weights[q];
    }

    auto is_fluid() const -> bool { return collision->is_fluid(); }
    auto operator[](std::size_t index) -> double& { return pdf[index]; }
    auto operator[](std::size_t index) const -> const double& { return pdf[index]; }

    // host-side single-cell collision: only NullCollision (a no-op) supports it
    auto collide(const uint_array<lattice_model::D>& lattice_position) -> void
    {
        collision->collide(*this, lattice_position);
    }

    auto density() const -> double { return collision->compute_density(*this); }
    auto velocity(double density) const -> double_array<lattice_model::D>
    {
        return collision->compute_velocity(*this, density);
    }
    auto equilibrium(double density, const double_array<lattice_model::D>& velocity) const
        -> double_array<lattice_model::Q>
    {
        return collision->compute_feq(density, velocity);
    }

    auto set_collision_handler(const Collision<lattice_model>* handler) -> void { collision = handler; }
    auto get_collision_handler() const -> decltype(collision) { return collision; }

    // true if any lattice neighbour (rest vector included) is an interior fluid cell (cell.hpp:75-92)
    auto has_fluid_vicinity(const Domain<lattice_model>& domain,
            const uint_array<lattice_model::D>& position) const -> bool;

    double* data() { return pdf.data(); }
    const double* data() const { return pdf.data(); }
};

} // namespace lbm

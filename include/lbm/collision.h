// include/lbm/collision.h -- collision operators of the B200 host surface.
//
// Class names, constructors and virtuals follow the reference's include/collision.h
// (:18-86).  The difference: an operator here is a DESCRIPTOR (kind + parameters)
// that Domain hands to the device through the C ABI (include/lbm_b200.h); the
// per-cell arithmetic of BGKCollision::collide (collision.hpp:61-70) runs inside
// the fused CUDA sweep, not on the host.  Calling collide() on the host throws:
// there is no CPU fallback.  A user-defined subclass with its own host collide()
// cannot run on the device and is rejected by Domain (device_kind() < 0).
#pragma once
#include <stdexcept>

#include "lbmdefinitions.h"
#include "../lbm_b200.h"

namespace lbm
{

template <typename lattice_model> class Cell;
template <typename lattice_model> class Domain;

template <typename lattice_model>
class Collision
{
public:
    // Moments / equilibrium of ONE host-side cell -- set-up and inspection helpers
    // (initial conditions via Cell::equilibrium, read-out of single cells).  Same
    // association as collision.hpp:7-51.  Never used inside the time loop.
    auto compute_density(const Cell<lattice_model>& cell) const -> double;
    auto compute_velocity(const Cell<lattice_model>& cell, double density) const
        -> double_array<lattice_model::D>;
    auto compute_feq(double density, const double_array<lattice_model::D>& velocity) const
        -> double_array<lattice_model::Q>;

    virtual bool is_fluid() const = 0;
    virtual void collide(Cell<lattice_model>& cell, const uint_array<lattice_model::D>& position) const = 0;
    virtual ~Collision() {}

    // --- device descriptor (new) ---
    // LBM_B200_* kind understood by the sweep, or -1 for operators that only exist as host code
    virtual int device_kind() const { return -1; }
    virtual lbm_b200_bc descriptor() const
    {
        lbm_b200_bc d{};
        d.kind = device_kind();
        d.rho = 1.0;
        return d;
    }

protected:
    [[noreturn]] static void host_collide_unsupported(const char* what)
    {
        throw std::logic_error(std::string(what) + "::collide runs inside the CUDA sweep "
                "(Domain::stream/swap/collide); there is no host implementation");
    }
};

template <typename lattice_model>
class FluidCollision : public Collision<lattice_model>
{
public:
    bool is_fluid() const override final { return true; }
};

template <typename lattice_model>
class NonFluidCollision : public Collision<lattice_model>
{
protected:
    Domain<lattice_model>& domain;

public:
    NonFluidCollision(Domain<lattice_model>& domain) : domain(domain) {}
    bool is_fluid() const override final { return false; }
};

template <typename lattice_model>
class BGKCollision : public FluidCollision<lattice_model>
{
    double tau { 0 };

public:
    explicit BGKCollision(double tau) : tau { tau } {}
    void collide(Cell<lattice_model>&, const uint_array<lattice_model::D>&) const override
    {
        this->host_collide_unsupported("BGKCollision");
    }
    int device_kind() const override { return LBM_B200_FLUID; }
    double relaxation_time() const { return tau; }
};

template <typename lattice_model>
class NullCollision : public Collision<lattice_model>
{
public:
    bool is_fluid() const override final { return false; }
    void collide(Cell<lattice_model>&, const uint_array<lattice_model::D>&) const override
    {
        // does nothing, on host and device alike (collision.h:81-85)
    }
    int device_kind() const override { return LBM_B200_NULL; }
};

} // namespace lbm

#include "cell.h"

namespace lbm
{

template <typename lattice_model>
inline double Collision<lattice_model>::compute_density(const Cell<lattice_model>& cell) const
{
    double density = 0;
    for (std::size_t q = 0; q < lattice_model::Q; ++q) density += cell[q];
    return density;
}

template <typename lattice_model>
inline auto Collision<lattice_model>::compute_velocity(const Cell<lattice_model>& cell, double density) const
    -> double_array<lattice_model::D>
{
    double_array<lattice_model::D> momentum = { 0.0, 0.0, 0.0 };
    for (std::size_t q = 0; q < lattice_model::Q; ++q)
        for (std::size_t d = 0; d < lattice_model::D; ++d)
            momentum[d] += cell[q] * lattice_model::velocities[q][d];
    for (std::size_t d = 0; d < lattice_model::D; ++d) momentum[d] /= density;
    return momentum;
}

template <typename lattice_model>
inline auto Collision<lattice_model>::compute_feq(double density,
        const double_array<lattice_model::D>& velocity) const -> double_array<lattice_model::Q>
{
    double_array<lattice_model::Q> feq;
    double uu = 0.0;
    for (std::size_t d = 0; d < lattice_model::D; ++d) uu += velocity[d] * velocity[d];
    const double uu_term = uu / (2 * C_S * C_S);
    for (std::size_t q = 0; q < lattice_model::Q; ++q) {
        double cu = 0.0;
        for (std::size_t d = 0; d < lattice_model::D; ++d) cu += lattice_model::velocities[q][d] * velocity[d];
        const double series = 1 + cu / (C_S * C_S) + cu * cu / (2 * C_S * C_S * C_S * C_S) - uu_term;
        feq[q] = lattice_model::weights[q] * density * series;
    }
    return feq;
}

} // namespace lbm

// include/lbm/boundary.h -- boundary handlers of the B200 host surface.
//
// Same classes and constructor signatures as the reference's include/boundary.h
// (:11-91).  Each handler is a parameter carrier; its arithmetic (boundary.hpp:
// 15-214) is evaluated link-wise inside the CUDA sweep (lbm_b200/csrc/kernels.cuh,
// link_value) and, when populations are read back, by the materialize kernel.
#pragma once
#include <list>
#include <memory>

#include "collision.h"

namespace lbm
{

#define LBM_B200_DEVICE_ONLY_COLLIDE(NAME)                                                        \
    void collide(Cell<lattice_model>&, const uint_array<lattice_model::D>&) const override      \
    {                                                                                             \
        this->host_collide_unsupported(NAME);                                                     \
    }

// half-way bounce-back wall (boundary.hpp:15-31)
template <typename lattice_model>
class NoSlipBoundary : public NonFluidCollision<lattice_model>
{
public:
    NoSlipBoundary(Domain<lattice_model>& domain) : NonFluidCollision<lattice_model>(domain) {}
    LBM_B200_DEVICE_ONLY_COLLIDE("NoSlipBoundary")
    int device_kind() const override { return LBM_B200_NOSLIP; }
};

// bounce-back with momentum injection 2 w_q rho (c_q . u_w) / c_s^2 (boundary.hpp:44-68)
template <typename lattice_model>
class MovingWallBoundary : public NonFluidCollision<lattice_model>
{
    double_array<lattice_model::D> wall_velocity;

public:
    MovingWallBoundary(Domain<lattice_model>& domain, const double_array<lattice_model::D>& wall_velocity)
        : NonFluidCollision<lattice_model>(domain), wall_velocity(wall_velocity) {}
    LBM_B200_DEVICE_ONLY_COLLIDE("MovingWallBoundary")
    int device_kind() const override { return LBM_B200_MOVINGWALL; }
    lbm_b200_bc descriptor() const override
    {
        lbm_b200_bc d = Collision<lattice_model>::descriptor();
        for (std::size_t k = 0; k < lattice_model::D; ++k) d.v[k] = wall_velocity[k];
        return d;
    }
};

// specular reflection (boundary.hpp:80-115)
template <typename lattice_model>
class FreeSlipBoundary : public NonFluidCollision<lattice_model>
{
public:
    FreeSlipBoundary(Domain<lattice_model>& domain) : NonFluidCollision<lattice_model>(domain) {}
    LBM_B200_DEVICE_ONLY_COLLIDE("FreeSlipBoundary")
    int device_kind() const override { return LBM_B200_FREESLIP; }
};

// anti-bounce-back against feq(rho_ref, u_neighbour) (boundary.hpp:129-150)
template <typename lattice_model>
class OutflowBoundary : public NonFluidCollision<lattice_model>
{
    double reference_density { 0.0 };

public:
    OutflowBoundary(Domain<lattice_model>& domain, double reference_density = 1.0)
        : NonFluidCollision<lattice_model>(domain), reference_density { reference_density } {}
    LBM_B200_DEVICE_ONLY_COLLIDE("OutflowBoundary")
    int device_kind() const override { return LBM_B200_OUTFLOW; }
    lbm_b200_bc descriptor() const override
    {
        lbm_b200_bc d = Collision<lattice_model>::descriptor();
        d.rho = reference_density;
        return d;
    }
};

// equilibrium inlet feq(rho_ref, u_in) (boundary.hpp:165-181)
template <typename lattice_model>
class InflowBoundary : public NonFluidCollision<lattice_model>
{
    double reference_density { 0.0 };
    double_array<lattice_model::D> inflow_velocity;

public:
    InflowBoundary(Domain<lattice_model>& domain, const double_array<lattice_model::D>& inflow_velocity,
            double reference_density = 1.0)
        : NonFluidCollision<lattice_model>(domain), reference_density { reference_density },
          inflow_velocity(inflow_velocity) {}
    LBM_B200_DEVICE_ONLY_COLLIDE("InflowBoundary")
    int device_kind() const override { return LBM_B200_INFLOW; }
    lbm_b200_bc descriptor() const override
    {
        lbm_b200_bc d = Collision<lattice_model>::descriptor();
        for (std::size_t k = 0; k < lattice_model::D; ++k) d.v[k] = inflow_velocity[k];
        d.rho = reference_density;
        return d;
    }
};

// like OutflowBoundary with a prescribed density (boundary.hpp:195-214)
template <typename lattice_model>
class PressureBoundary : public NonFluidCollision<lattice_model>
{
    double input_density;

public:
    PressureBoundary(Domain<lattice_model>& domain, double input_density)
        : NonFluidCollision<lattice_model>(domain), input_density { input_density } {}
    LBM_B200_DEVICE_ONLY_COLLIDE("PressureBoundary")
    int device_kind() const override { return LBM_B200_PRESSURE; }
    lbm_b200_bc descriptor() const override
    {
        lbm_b200_bc d = Collision<lattice_model>::descriptor();
        d.rho = input_density;
        return d;
    }
};

// Extension (not in the reference): a ghost-shell cell that mirrors its periodically
// wrapped interior image.  Equivalent to copying cell(wrapped) into the ghost cell
// before every stream() through the reference API.
template <typename lattice_model>
class PeriodicBoundary : public NonFluidCollision<lattice_model>
{
public:
    PeriodicBoundary(Domain<lattice_model>& domain) : NonFluidCollision<lattice_model>(domain) {}
    LBM_B200_DEVICE_ONLY_COLLIDE("PeriodicBoundary")
    int device_kind() const override { return LBM_B200_PERIODIC; }
};

#undef LBM_B200_DEVICE_ONLY_COLLIDE

// Owner of every boundary handler created while parsing a scenario; handlers live
// until program exit, as in boundary.h:73-91.
template <typename lattice_model>
class BoundaryKeeper
{
    static std::list<std::unique_ptr<NonFluidCollision<lattice_model>>>& store()
    {
        static std::list<std::unique_ptr<NonFluidCollision<lattice_model>>> handlers;
        return handlers;
    }

public:
    template <typename collision, typename... Args>
    static auto get_collision(Args&&... args) -> NonFluidCollision<lattice_model>&
    {
        store().emplace_back(new collision(args...));
        return *store().back();
    }
};

} // namespace lbm

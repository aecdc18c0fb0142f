// include/lbm/parallel.h -- the reference's parallel.h:11-23 declares a
// ParallelBoundary whose collide() is empty ("TODO: Not yet implemented") and an
// x-slab Domain::create_subdomain that nothing calls.  Here the decomposition is
// real: Domain splits the lattice into z-slabs across the GPUs of the box
// (lbm_b200_create_slab / lbm_b200_connect_local in include/lbm_b200.h) and the
// sweep kernel stores the populations that leave a slab directly into the
// neighbour's ghost plane over NVLink.  ParallelBoundary is kept as the no-op
// handler it is in the reference: cells tagged with it keep their stored values.
#pragma once
#include "collision.h"

namespace lbm
{
namespace parallel
{

template <typename lattice_model>
class ParallelBoundary : public NonFluidCollision<lattice_model>
{
public:
    ParallelBoundary(Domain<lattice_model>& domain) : NonFluidCollision<lattice_model>(domain) {}
    void collide(Cell<lattice_model>&, const uint_array<lattice_model::D>&) const override {}
    int device_kind() const override { return LBM_B200_PARALLEL; }
};

} // namespace parallel
} // namespace lbm

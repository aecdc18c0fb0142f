// oracle/ref_driver.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Drives the reference's OWN implementation of the hot path.  The headers
// model.h, lbmdefinitions.h, helper.h, collision.h/.hpp, cell.h/.hpp,
// domain.h/.hpp, boundary.h/.hpp and parallel.h are compiled, unmodified, from
// where they lie under /root/reference/include (see oracle/Makefile, target
// `ref`); nothing of them is copied into this repository.  Only the public API
// is used, in the order src/main.cpp:36-52 and io/scenario.h:91-188 use it:
//   BGKCollision<M>(tau); Domain<M>(xl,yl,zl,coll);
//   BoundaryKeeper<M>::get_collision<...>(domain, ...); setBoundaryCondition(...)
//   set_nonfluid_cells_nullcollide(); loop { stream(); swap(); collide(); }
// Output goes to oracle/_ref/libref_lbm.so (git-ignored, travels with gpurun).
#include <list>
#include <memory>
#include <vector>
#include <cstring>
#include <cstdio>
#include <stdexcept>

#include "model.h"
#include "parallel.h"
#include "lbmdefinitions.h"
#include "helper.h"
#include "collision.h"
#include "boundary.h"
#include "cell.h"
#include "domain.h"

#include "oracle.h"

namespace {

template <typename M>
struct Probe : lbm::FluidCollision<M> {
    // gives access to the non-virtual moment helpers of Collision<M>
    void collide(lbm::Cell<M>&, const lbm::uint_array<M::D>&) const override {}
};

template <typename M>
lbm::NonFluidCollision<M>& make_handler(lbm::Domain<M>& dom, const orc_box& b)
{
    using BK = lbm::BoundaryKeeper<M>;
    const lbm::double_array<M::D> v = { b.v[0], b.v[1], b.v[2] };
    switch (b.kind) {
    case ORC_NOSLIP:     return BK::template get_collision<lbm::NoSlipBoundary<M>>(dom);
    case ORC_MOVINGWALL: return BK::template get_collision<lbm::MovingWallBoundary<M>>(dom, v);
    case ORC_FREESLIP:   return BK::template get_collision<lbm::FreeSlipBoundary<M>>(dom);
    case ORC_OUTFLOW:    return BK::template get_collision<lbm::OutflowBoundary<M>>(dom, b.rho);
    case ORC_INFLOW:     return BK::template get_collision<lbm::InflowBoundary<M>>(dom, v, b.rho);
    case ORC_PRESSURE:   return BK::template get_collision<lbm::PressureBoundary<M>>(dom, b.rho);
    case ORC_PARALLEL:   return BK::template get_collision<lbm::parallel::ParallelBoundary<M>>(dom);
    default: throw std::logic_error("ref_driver: unknown boundary kind");
    }
}

template <typename M>
int kind_of(const lbm::Collision<M>* h)
{
    if (dynamic_cast<const lbm::FluidCollision<M>*>(h)) return ORC_FLUID;
    if (dynamic_cast<const lbm::NoSlipBoundary<M>*>(h)) return ORC_NOSLIP;
    if (dynamic_cast<const lbm::MovingWallBoundary<M>*>(h)) return ORC_MOVINGWALL;
    if (dynamic_cast<const lbm::FreeSlipBoundary<M>*>(h)) return ORC_FREESLIP;
    if (dynamic_cast<const lbm::OutflowBoundary<M>*>(h)) return ORC_OUTFLOW;
    if (dynamic_cast<const lbm::InflowBoundary<M>*>(h)) return ORC_INFLOW;
    if (dynamic_cast<const lbm::PressureBoundary<M>*>(h)) return ORC_PRESSURE;
    if (dynamic_cast<const lbm::NullCollision<M>*>(h)) return ORC_NULL;
    if (dynamic_cast<const lbm::parallel::ParallelBoundary<M>*>(h)) return ORC_PARALLEL;
    return -1;
}

template <typename M>
int run(const orc_case* c, orc_result* r)
{
    const int xl = (int) c->xl, yl = (int) c->yl, zl = (int) c->zl;
    omp_set_num_threads(c->threads > 0 ? c->threads : 1);

    auto collision = lbm::BGKCollision<M>(c->tau);
    lbm::Domain<M> domain(c->xl, c->yl, c->zl, collision);

    if (c->fluid_mask) {   // as io/vtk.hpp:137-150 does for a legacy-VTK mask
        auto& solid = lbm::BoundaryKeeper<M>::template get_collision<lbm::NoSlipBoundary<M>>(domain);
        size_t i = 0;
        for (int z = 1; z < zl + 1; ++z)
            for (int y = 1; y < yl + 1; ++y)
                for (int x = 1; x < xl + 1; ++x)
                    if (!c->fluid_mask[i++])
                        domain.cell(x, y, z).set_collision_handler(&solid);
        // io/vtk.hpp:145-146 tags the collide field only, so in the reference a
        // masked cell is solid/fluid on alternate steps (swap() exchanges the
        // vectors, is_fluid() is always asked of the collide field,
        // domain.hpp:74-83).  mask_literal keeps that; otherwise both fields are
        // tagged, as setBoundaryCondition (domain.hpp:189-190) does for boxes.
        if (!c->mask_literal) {
            domain.swap();
            i = 0;
            for (int z = 1; z < zl + 1; ++z)
                for (int y = 1; y < yl + 1; ++y)
                    for (int x = 1; x < xl + 1; ++x)
                        if (!c->fluid_mask[i++])
                            domain.cell(x, y, z).set_collision_handler(&solid);
            domain.swap();
        }
    }
    for (int b = 0; b < c->n_boxes; ++b) {
        const orc_box& box = c->boxes[b];
        domain.setBoundaryCondition(make_handler<M>(domain, box),
                box.x0, box.xE, box.y0, box.yE, box.z0, box.zE);
    }
    if (c->f_init) {
        for (int z = 0; z < zl + 2; ++z)
            for (int y = 0; y < yl + 2; ++y)
                for (int x = 0; x < xl + 2; ++x) {
                    const double* src = c->f_init + (size_t) domain.idx(x, y, z) * M::Q;
                    for (size_t q = 0; q < M::Q; ++q) domain.cell(x, y, z)[q] = src[q];
                }
    }
    if (c->null_opt) domain.set_nonfluid_cells_nullcollide();

    auto wrap = [](int v, int l) { return v < 1 ? v + l : (v > l ? v - l : v); };
    double seconds = 0.0;
    for (uint64_t t = 1; t <= c->steps; ++t) {
        if (c->periodic) {
            for (int z = 0; z < zl + 2; ++z)
                for (int y = 0; y < yl + 2; ++y)
                    for (int x = 0; x < xl + 2; ++x)
                        if (!domain.in_bounds(x, y, z))
                            domain.cell(x, y, z) = domain.cell(wrap(x, xl), wrap(y, yl), wrap(z, zl));
        }
        const double start = omp_get_wtime();
        domain.stream();
        domain.swap();
        domain.collide();
        if (t > c->untimed) seconds += omp_get_wtime() - start;
    }
    r->seconds = seconds;

    if (r->f || r->kind) {
        for (int z = 0; z < zl + 2; ++z)
            for (int y = 0; y < yl + 2; ++y)
                for (int x = 0; x < xl + 2; ++x) {
                    const size_t i = (size_t) domain.idx(x, y, z);
                    const auto& cell = domain.cell(x, y, z);
                    if (r->f) for (size_t q = 0; q < M::Q; ++q) r->f[i * M::Q + q] = cell[q];
                    if (r->kind) r->kind[i] = (uint8_t) kind_of<M>(cell.get_collision_handler());
                }
    }
    if (r->rho || r->u) {   // the loop of io/vtk.hpp:62-73
        size_t i = 0;
        for (int z = 1; z < zl + 1; ++z)
            for (int y = 1; y < yl + 1; ++y)
                for (int x = 1; x < xl + 1; ++x, ++i) {
                    auto current_cell = domain.cell(x, y, z);
                    auto density = current_cell.density();
                    auto vel = current_cell.velocity(density);
                    if (r->rho) r->rho[i] = density;
                    if (r->u) { r->u[3*i] = vel[0]; r->u[3*i+1] = vel[1]; r->u[3*i+2] = vel[2]; }
                }
    }
    return 0;
}

template <typename M>
lbm::Cell<M> cell_from(const Probe<M>& p, const double* f)
{
    lbm::Cell<M> cell(&p);
    for (size_t q = 0; q < M::Q; ++q) cell[q] = f[q];
    return cell;
}

#define DISPATCH(Q, EXPR15, EXPR19, EXPR27, BAD) \
    switch (Q) { case 15: { using M = lbm::model::d3q15; EXPR15; } break; \
                 case 19: { using M = lbm::model::d3q19; EXPR19; } break; \
                 case 27: { using M = lbm::model::d3q27; EXPR27; } break; \
                 default: BAD; }

template <typename M> double density_t(const double* f)
{ Probe<M> p; return cell_from<M>(p, f).density(); }
template <typename M> void velocity_t(const double* f, double rho, double* u)
{ Probe<M> p; auto v = cell_from<M>(p, f).velocity(rho); u[0] = v[0]; u[1] = v[1]; u[2] = v[2]; }
template <typename M> void feq_t(double rho, const double* u, double* out)
{ Probe<M> p; auto e = p.compute_feq(rho, { u[0], u[1], u[2] }); for (size_t q = 0; q < M::Q; ++q) out[q] = e[q]; }
template <typename M> void bgk_t(double tau, double* f)
{
    lbm::BGKCollision<M> bgk(tau);
    lbm::Cell<M> cell(&bgk);
    for (size_t q = 0; q < M::Q; ++q) cell[q] = f[q];
    cell.collide({ 1, 1, 1 });
    for (size_t q = 0; q < M::Q; ++q) f[q] = cell[q];
}
template <typename M> void model_t(double* vel, double* w)
{
    for (size_t q = 0; q < M::Q; ++q) {
        for (size_t d = 0; d < M::D; ++d) vel[q * 3 + d] = M::velocities[q][d];
        w[q] = M::weights[q];
    }
}

} // namespace

extern "C" {

int ref_run(const orc_case* c, orc_result* r)
{
    try {
        DISPATCH(c->Q, return run<M>(c, r), return run<M>(c, r), return run<M>(c, r), return -1)
    } catch (const std::exception& ex) {
        std::fprintf(stderr, "ref_run: %s\n", ex.what());
        return -2;
    }
    return -1;
}
double ref_density(int Q, const double* f)
{ DISPATCH(Q, return density_t<M>(f), return density_t<M>(f), return density_t<M>(f), return 0.0) return 0.0; }
void ref_velocity(int Q, const double* f, double rho, double* u)
{ DISPATCH(Q, velocity_t<M>(f, rho, u), velocity_t<M>(f, rho, u), velocity_t<M>(f, rho, u), return) }
void ref_feq(int Q, double rho, const double* u, double* out)
{ DISPATCH(Q, feq_t<M>(rho, u, out), feq_t<M>(rho, u, out), feq_t<M>(rho, u, out), return) }
void ref_bgk(int Q, double tau, double* f)
{ DISPATCH(Q, bgk_t<M>(tau, f), bgk_t<M>(tau, f), bgk_t<M>(tau, f), return) }
int ref_model(int Q, double* vel, double* w)
{ DISPATCH(Q, model_t<M>(vel, w), model_t<M>(vel, w), model_t<M>(vel, w), return -1) return 0; }
int ref_velocity_index(int Q, int u, int v, int w)
{
    DISPATCH(Q, return (int) M::velocity_index(u, v, w), return (int) M::velocity_index(u, v, w),
             return (int) M::velocity_index(u, v, w), return -1)
    return -1;
}

} // extern "C"

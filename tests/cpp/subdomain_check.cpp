// GPU check of Domain::create_subdomain (reference: domain.hpp:197-248), the x-interval helper of the
// reference's parallel.h sketch: new_xl = xend - xstart + 1, cells x in [xstart, xend) of every y, z copied to
// x - xstart + 1, the outer faces copied from the parent for the first / last rank, the cut faces tagged with the
// ParallelBoundary handler, origin shifted by rank * new_xl * spacing -- and the sub-domain is a working Domain.
//   usage: subdomain_check <scenario.cfg>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <memory>

#include "model.h"
#include "parallel.h"
#include "lbmdefinitions.h"
#include "helper.h"
#include "collision.h"
#include "boundary.h"
#include "cell.h"
#include "domain.h"
#include "io/configuration.h"
#include "io/vtk.h"
#include "io/scenario.h"

#define REQUIRE(cond) do { if (!(cond)) { std::printf("SUBDOMAIN_CHECK FAILED line %d: %s\n", __LINE__, #cond); return 1; } } while (0)

int main(int argc, char** argv)
{
    using model = lbm::model::d3q19;
    if (argc < 2) return 2;
    const char* cfg_argv[] = { "subdomain_check", argv[1] };
    lbm::io::Config cfg(2, const_cast<char**>(cfg_argv));
    auto collision = lbm::BGKCollision<model>(cfg.tau());
    auto domain = lbm::io::parse_scenario_file<model>(cfg.scenario_xml(), cfg, collision);
    for (int t = 0; t < 7; ++t) { domain->stream(); domain->swap(); domain->collide(); }

    const std::size_t xl = domain->xlength(), yl = domain->ylength(), zl = domain->zlength();
    lbm::parallel::ParallelBoundary<model> parallel_boundary(*domain);
    const int ranks = 2;
    const std::size_t half = xl / 2;
    for (int rank = 0; rank < ranks; ++rank) {
        const std::size_t xstart = rank == 0 ? 1 : half + 1, xend = rank == 0 ? half + 1 : xl + 1;
        auto sub = domain->create_subdomain(parallel_boundary, xstart, xend, rank, ranks);
        const std::size_t new_xl = xend - xstart + 1;                       // domain.hpp:202
        REQUIRE(sub->xlength() == new_xl && sub->ylength() == yl && sub->zlength() == zl);
        REQUIRE(sub->xorigin() == domain->xorigin() + rank * double(new_xl) * domain->xspacing());
        for (std::size_t z = 0; z < zl + 2; ++z)
            for (std::size_t y = 0; y < yl + 2; ++y) {
                for (std::size_t x = xstart; x < xend; ++x) {              // domain.hpp:208-214
                    const auto& a = sub->cell(int(x - xstart + 1), int(y), int(z));
                    const auto& b = domain->cell(int(x), int(y), int(z));
                    for (std::size_t q = 0; q < model::Q; ++q) REQUIRE(a[q] == b[q]);
                    REQUIRE(a.get_collision_handler()->device_kind() == b.get_collision_handler()->device_kind());
                }
                // cut faces carry the parallel handler, outer faces the parent's
                const auto* left = sub->cell(0, int(y), int(z)).get_collision_handler();
                const auto* right = sub->cell(int(new_xl + 1), int(y), int(z)).get_collision_handler();
                if (rank > 0) REQUIRE(left == &parallel_boundary);
                else REQUIRE(left->device_kind() == domain->cell(0, int(y), int(z)).get_collision_handler()->device_kind());
                if (rank < ranks - 1) REQUIRE(right == &parallel_boundary);
                else REQUIRE(right->device_kind() == domain->cell(int(xl + 1), int(y), int(z)).get_collision_handler()->device_kind());
            }
        // a sub-domain is a Domain: it steps (nothing is exchanged across the cut, like in the reference)
        sub->stream(); sub->swap(); sub->collide();
        REQUIRE(sub->timesteps_done() == 1);
        const auto& probe = sub->cell(1, 1, 1);
        REQUIRE(probe.density() > 0.5 && probe.density() < 1.5);
    }
    std::printf("SUBDOMAIN_CHECK OK\n");
    return 0;
}

// GPU check of the C++ host surface through the calls a user of the reference makes:
//   Config -> BGKCollision -> parse_scenario_file -> set_nonfluid_cells_nullcollide ->
//   cell(x,y,z)[q] = ... (initial condition) -> { stream(); swap(); collide(); } x N -> cell() read-out.
// Dumps the initial and final lattices; tests/test_host_surface_gpu.py compares with the oracle.
//   usage: host_api_check <lattice 15|19|27> <config.cfg> <steps> <out.bin>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <list>
#include <memory>
#include <vector>

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

template <typename model>
void dump(const lbm::Domain<model>& domain, std::FILE* out)
{
    const int xl = int(domain.xlength()), yl = int(domain.ylength()), zl = int(domain.zlength());
    std::vector<double> buf;
    for (int z = 0; z < zl + 2; ++z)
        for (int y = 0; y < yl + 2; ++y)
            for (int x = 0; x < xl + 2; ++x) {
                const auto& cell = domain.cell(x, y, z);
                for (std::size_t q = 0; q < model::Q; ++q) buf.push_back(cell[q]);
            }
    std::fwrite(buf.data(), sizeof(double), buf.size(), out);
}

template <typename model>
int run(int argc, char** argv)
{
    const char* cfg_argv[] = { "host_api_check", argv[2] };
    lbm::io::Config cfg(2, const_cast<char**>(cfg_argv));
    const int steps = std::atoi(argv[3]);
    auto collision = lbm::BGKCollision<model>(cfg.tau());
    auto domain = lbm::io::parse_scenario_file<model>(cfg.scenario_xml(), cfg, collision);
    domain->set_nonfluid_cells_nullcollide();

    // an initial condition written through the mutable Cell API, like user code of the reference would
    const int cx = int(domain->xlength()) / 2 + 1, cy = int(domain->ylength()) / 2 + 1, cz = int(domain->zlength()) / 2 + 1;
    auto& centre = domain->cell(cx, cy, cz);
    const auto feq = centre.equilibrium(1.03, { 0.02, -0.01, 0.015 });
    for (std::size_t q = 0; q < model::Q; ++q) centre[q] = feq[q];

    std::FILE* out = std::fopen(argv[4], "wb");
    if (!out) return 3;
    const int header[4] = { int(domain->xlength()), int(domain->ylength()), int(domain->zlength()), int(model::Q) };
    std::fwrite(header, sizeof(int), 4, out);
    dump(*domain, out);
    for (int t = 0; t < steps; ++t) {
        domain->stream();
        domain->swap();
        domain->collide();
    }
    dump(*domain, out);
    // handler kinds as the reference would report them through get_collision_handler()
    std::vector<unsigned char> kinds;
    for (int z = 0; z < header[2] + 2; ++z)
        for (int y = 0; y < header[1] + 2; ++y)
            for (int x = 0; x < header[0] + 2; ++x)
                kinds.push_back((unsigned char) domain->cell(x, y, z).get_collision_handler()->device_kind());
    std::fwrite(kinds.data(), 1, kinds.size(), out);
    // density / velocity through the device reduction
    std::vector<double> rho(std::size_t(header[0]) * header[1] * header[2]), u(3 * rho.size());
    domain->macroscopic(rho.data(), u.data());
    std::fwrite(rho.data(), sizeof(double), rho.size(), out);
    std::fwrite(u.data(), sizeof(double), u.size(), out);
    // ... and one .vts file
    lbm::io::write_vtk_file(*domain, cfg.output_dir(), cfg.output_filename(), steps);
    std::fclose(out);

    // misuse is reported, not silently computed on the host
    bool threw = false;
    try { domain->collide(); } catch (const std::logic_error&) { threw = true; }
    if (!threw) return 4;
    std::printf("HOST_API_CHECK DONE gpus=%zu\n", domain->gpu_count());
    return 0;
}

int main(int argc, char** argv)
{
    if (argc < 5) return 2;
    try {
        switch (std::atoi(argv[1])) {
        case 15: return run<lbm::model::d3q15>(argc, argv);
        case 27: return run<lbm::model::d3q27>(argc, argv);
        default: return run<lbm::model::d3q19>(argc, argv);
        }
    } catch (const std::exception& ex) {
        std::cerr << "An error occured: " << ex.what() << std::endl;
        return 1;
    }
}

// src/main.cpp -- driver of the B200 build.
//
// Same flow as the reference's src/main.cpp:30-73 (config -> BGK operator ->
// scenario -> time loop -> MLUPS), with two differences: the velocity set is
// chosen at run time (`lattice = 15|19|27` in the config) instead of by the
// compile-time `#define D3Q`, and time is measured on the device around the
// whole loop so that steps are queued back to back.  The reference's unmodified
// src/main.cpp also compiles against include/lbm (tests/test_host_surface.py).
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

template <typename model>
int simulate(lbm::io::Config& cfg)
{
    lbm::device::set_gpus(int(cfg.gpus()));
    lbm::device::set_arithmetic(cfg.arithmetic() == "exact" ? LBM_B200_EXACT : LBM_B200_FAST);

    auto collision = lbm::BGKCollision<model>(cfg.tau());
    auto domain = lbm::io::parse_scenario_file<model>(cfg.scenario_xml(), cfg, collision);
    std::cout << cfg << std::endl;
    std::cout << "> Lattice:                " << model::name << '\n'
              << "> GPUs (" << domain->split_axis_name() << "-slabs):         " << domain->gpu_count() << '\n'
              << "> Arithmetic:             " << cfg.arithmetic() << std::endl;
    std::cout << "> Domain lengths (x,y,z): " << domain->xlength() << ", " << domain->ylength() << ", "
              << domain->zlength() << std::endl;
    std::cout << "Starting simulation..." << std::endl;

    domain->set_nonfluid_cells_nullcollide();
    domain->step(0);                     // pushes geometry before the clock starts
    domain->synchronize();

    double sweep_seconds = 0.0;
    const double walltime = omp_get_wtime();
    const auto plot_every = cfg.timesteps_per_plot();
    std::uint64_t t = 0;
    while (t < cfg.timesteps()) {
        // run up to the next output step in one batch; the device queue stays full
        std::uint64_t batch = cfg.timesteps() - t;
        if (plot_every) batch = std::min<std::uint64_t>(batch, plot_every - t % plot_every);
        const double start = omp_get_wtime();
        domain->step(batch);
        domain->synchronize();
        sweep_seconds += omp_get_wtime() - start;
        t += batch;
        if (plot_every && t % plot_every == 0)
            lbm::io::write_vtk_file(*domain, cfg.output_dir(), cfg.output_filename(), t);
        std::cout << "\r" << int(double(t) / cfg.timesteps() * 100) << " %";
        std::cout.flush();
    }
    std::cout << "\nFinished!" << std::endl;
    std::cout << "Total runtime: " << omp_get_wtime() - walltime << " seconds." << std::endl;

    const double steps = double(cfg.timesteps());
    const double all_cells = double(2 + domain->xlength()) * (2 + domain->ylength()) * (2 + domain->zlength());
    const double interior = double(domain->xlength()) * domain->ylength() * domain->zlength();
    // the reference's figure counts ghost cells (src/main.cpp:64-65); the second one counts updates
    std::cout << "MLUPS: " << all_cells * steps / (sweep_seconds * 1e6) << std::endl;
    std::cout << "MLUPS (interior cell updates): " << interior * steps / (sweep_seconds * 1e6) << std::endl;
    std::cout << "Effective bandwidth: " << interior * steps * 2 * model::Q * 8 / sweep_seconds / 1e9
              << " GB/s (2*Q*8 bytes per update)" << std::endl;
    return EXIT_SUCCESS;
}

int main(int argc, char** argv)
{
    try {
        lbm::io::Config cfg(argc, argv);
        switch (cfg.lattice()) {
        case 15: return simulate<lbm::model::d3q15>(cfg);
        case 27: return simulate<lbm::model::d3q27>(cfg);
        default: return simulate<lbm::model::d3q19>(cfg);
        }
    } catch (const std::exception& ex) {
        std::cerr << "An error occured: " << ex.what() << std::endl;
        return EXIT_FAILURE;
    }
}

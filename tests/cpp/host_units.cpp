// Host-only unit checks of the C++ surface (no GPU needed): descriptors, single-cell helpers against the
// oracle, config-file / command-line parsing, scenario-XML parsing, error behaviour.
// Built and run by tests/test_host_surface.py; prints HOST_UNITS OK on success.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <list>
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

#include "oracle.h"   // test infrastructure (tests may use the checker; the product may not)

static int failures = 0;
#define CHECK(cond) do { if (!(cond)) { std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); ++failures; } } while (0)

template <typename M>
void check_model()
{
    double vel[27 * 3], w[27], ovel[27 * 3], ow[27];
    CHECK(lbm_b200_model(int(M::Q), vel, w) == 0);
    CHECK(oracle_model(int(M::Q), ovel, ow) == 0);
    for (std::size_t q = 0; q < M::Q; ++q) {
        for (std::size_t d = 0; d < M::D; ++d) {
            CHECK(M::velocities[q][d] == vel[3 * q + d]);
            CHECK(M::velocities[q][d] == ovel[3 * q + d]);
        }
        CHECK(std::memcmp(&M::weights[q], &ow[q], 8) == 0);
        CHECK(M::weights[q] == w[q]);
        CHECK(M::velocity_index(int(vel[3 * q]), int(vel[3 * q + 1]), int(vel[3 * q + 2])) == q);
        CHECK(int(M::velocity_index(int(vel[3 * q]), int(vel[3 * q + 1]), int(vel[3 * q + 2])))
                == oracle_velocity_index(int(M::Q), int(vel[3 * q]), int(vel[3 * q + 1]), int(vel[3 * q + 2])));
        const int qi = M::inv(int(q));
        for (std::size_t d = 0; d < M::D; ++d) CHECK(M::velocities[qi][d] == -M::velocities[q][d]);
    }
    // single-cell helpers (Cell::density/velocity/equilibrium) are bit-identical to the oracle
    lbm::BGKCollision<M> bgk(0.7);
    lbm::Cell<M> cell(&bgk);
    CHECK(cell.is_fluid());
    for (std::size_t q = 0; q < M::Q; ++q) {
        CHECK(cell[q] == M::weights[q]);                  // cell.hpp:9-15
        cell[q] = M::weights[q] * (1.0 + 0.1 * std::sin(1.0 + q));
    }
    const double rho = cell.density();
    CHECK(rho == oracle_density(int(M::Q), cell.data()));
    const auto u = cell.velocity(rho);
    double ou[3];
    oracle_velocity(int(M::Q), cell.data(), rho, ou);
    CHECK(std::memcmp(u.data(), ou, sizeof ou) == 0);
    const auto feq = cell.equilibrium(rho, u);
    double ofeq[27];
    oracle_feq(int(M::Q), rho, ou, ofeq);
    CHECK(std::memcmp(feq.data(), ofeq, M::Q * sizeof(double)) == 0);
    // host-side collide does not exist
    bool threw = false;
    try { cell.collide({ 1, 1, 1 }); } catch (const std::logic_error&) { threw = true; }
    CHECK(threw);
}

static void check_config(const char* tmpdir)
{
    const std::string cfgfile = std::string(tmpdir) + "/t.cfg";
    const std::string outdir = std::string(tmpdir) + "/out";
    {
        std::ofstream f(cfgfile);
        f << "# comment\ncollision-model = bgk\ntau             = 0.6\ntimesteps       = 1000\n"
             "timesteps-per-plot = 1   # trailing comment\noutput-dir      = " << outdir << "\nscenario-file   = scenarios/cavity64.xml\n";
    }
    {
        const char* argv[] = { "lbm", cfgfile.c_str() };
        lbm::io::Config cfg(2, const_cast<char**>(argv));
        CHECK(cfg.tau() == 0.6);
        CHECK(cfg.timesteps() == 1000);
        CHECK(cfg.timesteps_per_plot() == 1);
        CHECK(cfg.collision_model() == "bgk");
        CHECK(cfg.scenario_xml() == "scenarios/cavity64.xml");
        CHECK(cfg.omp_threads() == 1);          // io/configuration.h:24 default
        CHECK(cfg.output_filename() == "output");
        CHECK(cfg.output_dir() == outdir);
        CHECK(cfg.input_file() == cfgfile);
        CHECK(file_exists(outdir));
        std::ostringstream echo;
        echo << cfg;
        CHECK(echo.str().find("> Tau:                    0.6") != std::string::npos);
    }
    {   // command line overrides the file; unknown options are ignored; -t short form; --key=value form
        std::ofstream(outdir + "/old.7.vts") << "x";
        std::ofstream(outdir + "/keep.txt") << "x";
        const char* argv[] = { "lbm", "--unknown-flag", "-t", "25", "--tau=0.8", cfgfile.c_str(), "--gpus", "2", "--lattice", "27" };
        lbm::io::Config cfg(10, const_cast<char**>(argv));
        CHECK(cfg.timesteps() == 25);
        CHECK(cfg.tau() == 0.8);
        CHECK(cfg.gpus() == 2 && cfg.lattice() == 27);
        CHECK(!file_exists(outdir + "/old.7.vts"));   // earlier plots are removed ...
        CHECK(file_exists(outdir + "/keep.txt"));     // ... other files are not (deviation from io/configuration.h:138-143)
    }
    auto expect_throw = [&](std::vector<const char*> args) {
        bool threw = false;
        try { lbm::io::Config cfg(int(args.size()), const_cast<char**>(args.data())); } catch (const std::exception&) { threw = true; }
        CHECK(threw);
    };
    expect_throw({ "lbm", cfgfile.c_str(), "--tau", "0.4" });        // (0.5, 2.0)
    expect_throw({ "lbm", cfgfile.c_str(), "--tau", "abc" });
    expect_throw({ "lbm", cfgfile.c_str(), "--collision-model", "mrt" });
    expect_throw({ "lbm", "--tau", "0.6", "--timesteps", "3" });     // required keys missing
    expect_throw({ "lbm", "/nonexistent/file.cfg" });
}

static void check_xml(const char* tmpdir)
{
    using namespace lbm::io;
    const std::string file = std::string(tmpdir) + "/s.xml";
    {
        std::ofstream f(file);
        f << "<?xml version=\"1.0\" ?>\n<!-- c --><scenario name=\"A &amp; B\">\n  <!-- <domain vtk-file=\"\"> -->\n"
             "  <domain xl=\"4\" yl='5' zl=\"6\">\n    <boundary extent=\"z0\" condition=\"noslip\" />\n"
             "    <boundary extent=\"0 5 6 6 0 7\" condition=\"movingwall\" vx=\"0.05\" vy=\"0\" vz=\"-1e-2\"/>\n  </domain>\n</scenario>\n";
    }
    xml::Node doc;
    CHECK(xml::load_file(file, doc));
    const xml::Node* sc = doc.child("scenario");
    CHECK(sc && *sc->attribute("name") == "A & B");
    const xml::Node* dom = sc->child("domain");
    CHECK(dom && xml::as_uint(*dom->attribute("yl")) == 5 && dom->children.size() == 2);
    CHECK(xml::as_double(*dom->children[1].attribute("vz")) == -1e-2);
    CHECK(!dom->attribute("vtk-file"));
    CHECK(!xml::load_file(std::string(tmpdir) + "/missing.xml", doc));
    { std::ofstream f(file); f << "<scenario name=\"x\"><domain></scenario>"; }
    CHECK(!xml::load_file(file, doc));
    // shipped scenarios parse
    for (const char* s : { "scenarios/cavity64.xml", "scenarios/cavity512.xml", "scenarios/channel_d3q27.xml",
                           "scenarios/step_small.xml", "scenarios/shear_small.xml" })
        CHECK(xml::load_file(s, doc) && doc.child("scenario") && doc.child("scenario")->child("domain"));
}

static void check_scenario_errors(const char* tmpdir)
{
    using M = lbm::model::d3q19;
    const std::string cfgfile = std::string(tmpdir) + "/t.cfg";
    const char* argv[] = { "lbm", cfgfile.c_str() };
    lbm::io::Config cfg(2, const_cast<char**>(argv));
    lbm::BGKCollision<M> bgk(0.6);
    auto expect = [&](const std::string& xml_text, const std::string& needle) {
        const std::string file = std::string(tmpdir) + "/bad.xml";
        { std::ofstream f(file); f << xml_text; }
        std::string what;
        try { lbm::io::parse_scenario_file<M>(file, cfg, bgk); } catch (const std::logic_error& e) { what = e.what(); }
        if (what.find(needle) == std::string::npos) { std::printf("FAILED: expected \"%s\" in \"%s\"\n", needle.c_str(), what.c_str()); ++failures; }
    };
    // messages of io/scenario.h:138-175
    expect("<nonsense", "could not be read properly!");
    expect("<other name=\"x\"/>", "Scenario node missing!");
    expect("<scenario><domain xl=\"2\" yl=\"2\" zl=\"2\"/></scenario>", "Scenario name is missing!");
    expect("<scenario name=\"s\"></scenario>", "Domain node is missing!");
    expect("<scenario name=\"s\"><domain xl=\"2\" yl=\"2\"/></scenario>", "Neither vtk-file nor xl/yl/zl attribute provided to domain node!");
    expect("<scenario name=\"s\"><domain vtk-file=\"/nonexistent.vtk\"/></scenario>", "does not exist or does not seem to be a valid structured grids file!");
    std::string what;
    try { lbm::io::parse_scenario_file<M>(std::string(tmpdir) + "/absent.xml", cfg, bgk); } catch (const std::logic_error& e) { what = e.what(); }
    CHECK(what.find("could not be read properly!") != std::string::npos);
    CHECK(cfg.output_filename() == "s");      // set from the scenario name before the domain is built (io/scenario.h:149)
}

int main(int argc, char** argv)
{
    if (argc < 2) { std::printf("usage: host_units <tmpdir>\n"); return 2; }
    check_model<lbm::model::d3q15>();
    check_model<lbm::model::d3q19>();
    check_model<lbm::model::d3q27>();
    CHECK(lbm::C_S * lbm::C_S == 0.33333333333376547);
    check_config(argv[1]);
    check_xml(argv[1]);
    check_scenario_errors(argv[1]);
    if (lbm_b200_device_count() == 0) {   // without a GPU a Domain cannot exist: loud failure, no fallback
        bool threw = false;
        lbm::BGKCollision<lbm::model::d3q19> bgk(0.6);
        try { lbm::Domain<lbm::model::d3q19> d(4, 4, 4, bgk); } catch (const std::runtime_error&) { threw = true; }
        CHECK(threw);
    }
    std::printf(failures ? "HOST_UNITS FAILED (%d)\n" : "HOST_UNITS OK\n", failures);
    return failures ? 1 : 0;
}

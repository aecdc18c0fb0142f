// include/lbm/io/configuration.h -- run configuration (command line + key = value file).
//
// Same class, accessors, option names, required keys and echo format as the
// reference's include/io/configuration.h:14-161, which is built on
// boost::program_options / boost::filesystem.  Boost is not a dependency here:
// the small subset the reference uses (long/short options, one positional input
// file, `key = value` lines with # comments, unknown options ignored on the
// command line) is parsed directly.  Extra keys for the GPU build: `gpus`,
// `arithmetic` (fast | exact), `lattice` (15 | 19 | 27).
// Deviations, on purpose: range violations throw std::invalid_argument instead of
// assert()-aborting (io/configuration.h:123-128), and an existing output directory
// is only cleared of *.vts files rather than of every file (:138-143).
#pragma once
#include <cstdint>
#include <cstdlib>
#include <dirent.h>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <sys/stat.h>
#include <unistd.h>
#include <vector>

#if defined(_OPENMP)
#include <omp.h>
#else
#include <chrono>
// the driver (src/main.cpp:47-63) times with omp_get_wtime(); provide it when OpenMP is off
inline double omp_get_wtime()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
inline int omp_get_max_threads() { return 1; }
inline void omp_set_num_threads(int) {}
#endif

namespace lbm
{
namespace io
{

class Config
{
    std::string _collision_model { "bgk" };
    std::string _input_file;
    std::string _output_dir { "vtk" };
    std::string _output_filename { "output" };
    std::string _scenario_xml;
    std::uint64_t _timesteps { 0 };
    std::uint64_t _timesteps_per_plot { 0 };
    double _tau { 1.0 };
    std::uint32_t _omp_threads { 1 };
    std::uint32_t _gpus { 1 };
    std::uint32_t _lattice { 19 };
    std::string _arithmetic { "fast" };

    struct Option {
        const char* name;
        char short_name;
        bool takes_value;
        const char* help;
    };
    static const std::vector<Option>& options()
    {
        static const std::vector<Option> opts = {
            { "help", 'h', false, "Show help message" },
            { "input-file", 'i', true, "Input file containing configuration values" },
            { "collision-model", 'c', true, "Collision operator model" },
            { "tau", 0, true, "Relaxation factor for BGK collision operator. Must be in (0.5, 2.0)" },
            { "timesteps", 't', true, "Number of steps to perform" },
            { "timesteps-per-plot", 0, true, "Number of timesteps after which an output file is written" },
            { "scenario-file", 0, true, "XML file containing scenario to simulate" },
            { "omp-threads", 0, true, "Number of OpenMP threads to use (host-side set-up only)" },
            { "output-dir", 0, true, "Output directory for plots" },
            { "gpus", 0, true, "Number of GPUs (z-slabs)" },
            { "arithmetic", 0, true, "fast (reciprocals + FMA) or exact (bit-identical to the CPU build)" },
            { "lattice", 0, true, "Velocity set: 15, 19 or 27" },
        };
        return opts;
    }

    static std::string trim(const std::string& s)
    {
        const auto b = s.find_first_not_of(" \t\r\n");
        if (b == std::string::npos) return "";
        const auto e = s.find_last_not_of(" \t\r\n");
        return s.substr(b, e - b + 1);
    }

    template <typename T>
    static T convert(const std::string& key, const std::string& text)
    {
        std::istringstream in(text);
        T value {};
        in >> value;
        if (in.fail() || !(in >> std::ws).eof())
            throw std::invalid_argument("the argument ('" + text + "') for option '--" + key + "' is invalid");
        return value;
    }

    void assign(const std::string& key, const std::string& value, std::map<std::string, bool>& seen, bool override_existing)
    {
        if (seen[key] && !override_existing) return;   // command line wins over the file, as with po::store order
        seen[key] = true;
        if (key == "input-file") _input_file = value;
        else if (key == "collision-model") _collision_model = value;
        else if (key == "tau") _tau = convert<double>(key, value);
        else if (key == "timesteps") _timesteps = convert<std::uint64_t>(key, value);
        else if (key == "timesteps-per-plot") _timesteps_per_plot = convert<std::uint64_t>(key, value);
        else if (key == "scenario-file") _scenario_xml = value;
        else if (key == "omp-threads") _omp_threads = convert<std::uint32_t>(key, value);
        else if (key == "output-dir") _output_dir = value;
        else if (key == "gpus") _gpus = convert<std::uint32_t>(key, value);
        else if (key == "arithmetic") _arithmetic = value;
        else if (key == "lattice") _lattice = convert<std::uint32_t>(key, value);
    }

    static const Option* find_long(const std::string& name)
    {
        for (const auto& o : options())
            if (name == o.name) return &o;
        return nullptr;
    }
    static const Option* find_short(char c)
    {
        for (const auto& o : options())
            if (o.short_name && c == o.short_name) return &o;
        return nullptr;
    }

    static void print_usage()
    {
        std::cout << "Usage: lbm_isotropy [file] additional-options..." << std::endl;
        for (const auto& o : options()) {
            std::cout << "  ";
            if (o.short_name) std::cout << "-" << o.short_name << " [ --" << o.name << " ]";
            else std::cout << "--" << o.name;
            if (o.takes_value) std::cout << " arg";
            std::cout << "\n        " << o.help << "\n";
        }
        std::cout << std::endl;
    }

    void parse_file(std::map<std::string, bool>& seen)
    {
        std::ifstream in(_input_file);
        if (!in) throw std::invalid_argument("can not read options configuration file '" + _input_file + "'");
        std::string line;
        while (std::getline(in, line)) {
            const auto hash = line.find('#');
            if (hash != std::string::npos) line.erase(hash);
            line = trim(line);
            if (line.empty()) continue;
            const auto eq = line.find('=');
            if (eq == std::string::npos) throw std::invalid_argument("the options configuration file contains an invalid line '" + line + "'");
            const std::string key = trim(line.substr(0, eq));
            if (!find_long(key) || key == "help" || key == "input-file")
                throw std::invalid_argument("unrecognised option '" + key + "'");
            assign(key, trim(line.substr(eq + 1)), seen, false);
        }
    }

public:
    auto collision_model() const -> decltype(_collision_model) { return _collision_model; }
    auto input_file() const -> decltype(_input_file) { return _input_file; }
    auto output_dir() const -> decltype(_output_dir) { return _output_dir; }
    auto output_filename() const -> decltype(_output_filename) { return _output_filename; }
    void set_output_filename(const std::string& filename) { _output_filename = filename; }
    auto timesteps() const -> decltype(_timesteps) { return _timesteps; }
    auto timesteps_per_plot() const -> decltype(_timesteps_per_plot) { return _timesteps_per_plot; }
    auto tau() const -> decltype(_tau) { return _tau; }
    auto omp_threads() const -> decltype(_omp_threads) { return _omp_threads; }
    auto scenario_xml() const -> decltype(_scenario_xml) { return _scenario_xml; }
    auto gpus() const -> decltype(_gpus) { return _gpus; }
    auto lattice() const -> decltype(_lattice) { return _lattice; }
    auto arithmetic() const -> decltype(_arithmetic) { return _arithmetic; }

    Config(int argc, char** argv)
    {
        std::map<std::string, bool> seen;
        bool help = false;
        bool have_positional = false;
        for (int i = 1; i < argc; ++i) {
            const std::string arg = argv[i];
            const Option* opt = nullptr;
            std::string value;
            bool have_value = false;
            if (arg.size() > 2 && arg[0] == '-' && arg[1] == '-') {
                std::string name = arg.substr(2);
                const auto eq = name.find('=');
                if (eq != std::string::npos) {
                    value = name.substr(eq + 1);
                    name.erase(eq);
                    have_value = true;
                }
                opt = find_long(name);
                if (!opt) continue;                       // allow_unregistered()
            } else if (arg.size() >= 2 && arg[0] == '-' && arg[1] != '-') {
                opt = find_short(arg[1]);
                if (!opt) continue;
                if (arg.size() > 2) {
                    value = arg.substr(2);
                    have_value = true;
                }
            } else {
                if (!have_positional) {                   // positional: one file name
                    assign("input-file", arg, seen, true);
                    have_positional = true;
                }
                continue;
            }
            if (!opt->takes_value) {
                if (std::string(opt->name) == "help") help = true;
                continue;
            }
            if (!have_value) {
                if (i + 1 >= argc) throw std::invalid_argument(std::string("the required argument for option '--") + opt->name + "' is missing");
                value = argv[++i];
            }
            assign(opt->name, value, seen, true);
        }

        std::cout << "LBM simulation by Krivokapic, Mody, Malcher" << std::endl;
        if (help || argc == 1) {
            print_usage();
            std::exit(1);
        }
        if (!_input_file.empty()) {
            std::cout << "Reading configuration file..." << std::endl;
            parse_file(seen);
        }
        for (const char* required : { "tau", "timesteps", "timesteps-per-plot", "scenario-file" })
            if (!seen[required])
                throw std::invalid_argument(std::string("the option '--") + required + "' is required but missing");

        if (!(_timesteps > 0)) throw std::invalid_argument("timesteps must be positive");
        if (!(_tau > 0.5 && _tau < 2.0)) throw std::invalid_argument("tau must be in (0.5, 2.0)");
        if (!(_omp_threads > 0)) throw std::invalid_argument("omp-threads must be positive");
        if (_collision_model != "bgk") throw std::invalid_argument("only the bgk collision model is supported");
        if (_lattice != 15 && _lattice != 19 && _lattice != 27) throw std::invalid_argument("lattice must be 15, 19 or 27");
        if (_arithmetic != "fast" && _arithmetic != "exact") throw std::invalid_argument("arithmetic must be fast or exact");
        if (_gpus < 1) throw std::invalid_argument("gpus must be positive");
        omp_set_num_threads(int(_omp_threads));

        struct stat st;
        if (stat(_output_dir.c_str(), &st) != 0) {
            std::cout << "Output directory \"" + _output_dir + "\" does not exist. Creating." << std::endl;
            if (mkdir(_output_dir.c_str(), 0777) != 0)
                throw std::runtime_error("cannot create output directory \"" + _output_dir + "\"");
        } else if (DIR* dir = opendir(_output_dir.c_str())) {
            // clean files of earlier runs
            while (dirent* entry = readdir(dir)) {
                const std::string name = entry->d_name;
                if (name.size() > 4 && name.compare(name.size() - 4, 4, ".vts") == 0)
                    unlink((_output_dir + "/" + name).c_str());
            }
            closedir(dir);
        }
    }
};

inline auto operator<<(std::ostream& lhs, const lbm::io::Config& cfg) -> decltype(lhs)
{
    lhs << "Configuration:" << '\n'
        << "> Configuration file:     " << cfg.input_file() << '\n'
        << "> Output directory:       " << cfg.output_dir() << '\n'
        << "> Output filename:        " << cfg.output_filename() << '\n'
        << "> Number of OMP threads:  " << omp_get_max_threads() << '\n'
        << "> Collision model:        " << cfg.collision_model() << '\n'
        << "> Tau:                    " << cfg.tau() << '\n'
        << "> Timesteps:              " << cfg.timesteps() << '\n'
        << "> Timesteps per plot:     " << cfg.timesteps_per_plot() << '\n'
        << "> Scenario file:          " << cfg.scenario_xml();
    return lhs;
}

} // namespace io
} // namespace lbm

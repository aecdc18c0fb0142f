// include/lbm/io/vtk.h -- VTK output and geometry input without the VTK library.
//
//   write_vtk_file       same signature, file name (<dir>/<name>.<t>.vts), point order
//                        (z,y,x over interior cells), point coordinates
//                        (spacing*i + origin - 1) and array names ("Velocity",
//                        "Density") as the reference's io/vtk.hpp:20-89.  Density and
//                        velocity are reduced ON THE DEVICE (north star item 5) and only
//                        those 4 doubles per cell cross PCIe.  The file is a VTK XML
//                        StructuredGrid with raw appended Float64 data (ASCII when
//                        LBM_VTS_ASCII is defined) -- readable by ParaView like the
//                        reference's output.
//   read_vtk_point_file  legacy-VTK ASCII STRUCTURED_POINTS fluid mask -> Domain with
//                        solid cells tagged, as io/vtk.hpp:94-157.  The reference tags the
//                        collide field only (vtk.hpp:145-146), which makes masked cells flip
//                        between solid and fluid on every swap(); that behaviour is opt-in
//                        (lbm::device::set_mask_literal / LBM_B200_MASK_LITERAL=1, bit-identical
//                        to the reference); by default a masked cell is solid in both lattices.
#pragma once
#include <cctype>
#include <cstdint>
#include <fstream>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "domain.h"
#include "boundary.h"
#include "helper.h"

namespace lbm
{
namespace io
{

template <typename lattice_model>
void write_vtk_file(const Domain<lattice_model>& domain, const std::string& output_dir,
        const std::string& output_filename, std::uint64_t t)
{
    const auto xl = domain.xlength(), yl = domain.ylength(), zl = domain.zlength();
    const std::size_t n = xl * yl * zl;
    std::vector<double> density(n), velocity(3 * n);
    domain.macroscopic(density.data(), velocity.data());

    std::vector<double> points(3 * n);
    std::size_t i = 0;
    for (std::size_t z = 1; z < zl + 1; ++z)
        for (std::size_t y = 1; y < yl + 1; ++y)
            for (std::size_t x = 1; x < xl + 1; ++x, ++i) {
                points[3 * i + 0] = domain.xspacing() * x + domain.xorigin() - 1;
                points[3 * i + 1] = domain.yspacing() * y + domain.yorigin() - 1;
                points[3 * i + 2] = domain.zspacing() * z + domain.zorigin() - 1;
            }

    std::stringstream name;
    name << output_dir << "/" << output_filename << "." << t << ".vts";
    std::ofstream file(name.str(), std::ios::trunc | std::ios::binary);
    if (!file) throw std::runtime_error("cannot open \"" + name.str() + "\" for writing");

    std::ostringstream extent;
    extent << "0 " << xl - 1 << " 0 " << yl - 1 << " 0 " << zl - 1;
    file << "<?xml version=\"1.0\"?>\n"
         << "<VTKFile type=\"StructuredGrid\" version=\"1.0\" byte_order=\"LittleEndian\" header_type=\"UInt64\">\n"
         << "  <StructuredGrid WholeExtent=\"" << extent.str() << "\">\n"
         << "    <Piece Extent=\"" << extent.str() << "\">\n"
         << "      <PointData Scalars=\"Density\" Vectors=\"Velocity\">\n";
#if defined(LBM_VTS_ASCII)
    file.precision(17);
    auto ascii = [&](const char* label, int comps, const std::vector<double>& data) {
        file << "        <DataArray type=\"Float64\" Name=\"" << label << "\" NumberOfComponents=\"" << comps
             << "\" format=\"ascii\">\n";
        for (std::size_t k = 0; k < data.size(); ++k) file << data[k] << ((k + 1) % comps ? ' ' : '\n');
        file << "        </DataArray>\n";
    };
    ascii("Velocity", 3, velocity);
    ascii("Density", 1, density);
    file << "      </PointData>\n      <Points>\n";
    ascii("Points", 3, points);
    file << "      </Points>\n    </Piece>\n  </StructuredGrid>\n</VTKFile>\n";
#else
    const std::uint64_t bytes_v = velocity.size() * sizeof(double), bytes_d = density.size() * sizeof(double);
    file << "        <DataArray type=\"Float64\" Name=\"Velocity\" NumberOfComponents=\"3\" format=\"appended\" offset=\"0\"/>\n"
         << "        <DataArray type=\"Float64\" Name=\"Density\" NumberOfComponents=\"1\" format=\"appended\" offset=\""
         << bytes_v + 8 << "\"/>\n"
         << "      </PointData>\n      <Points>\n"
         << "        <DataArray type=\"Float64\" Name=\"Points\" NumberOfComponents=\"3\" format=\"appended\" offset=\""
         << bytes_v + 8 + bytes_d + 8 << "\"/>\n"
         << "      </Points>\n    </Piece>\n  </StructuredGrid>\n"
         << "  <AppendedData encoding=\"raw\">\n_";
    auto block = [&](const std::vector<double>& data) {
        const std::uint64_t bytes = data.size() * sizeof(double);
        file.write(reinterpret_cast<const char*>(&bytes), sizeof bytes);
        file.write(reinterpret_cast<const char*>(data.data()), std::streamsize(bytes));
    };
    block(velocity);
    block(density);
    block(points);
    file << "\n  </AppendedData>\n</VTKFile>\n";
#endif
    file.close();
    if (!file) throw std::runtime_error("failed to write \"" + name.str() + "\"");
}

template <typename lattice_model, typename solid_collision_model>
auto read_vtk_point_file(const std::string& filename, FluidCollision<lattice_model>& fluid_collision_model)
    -> Domain_ptr<lattice_model>
{
    std::ifstream in(filename);
    if (!in)
        throw std::logic_error("VTK file \"" + filename + "\" does not exist or does not "
                "seem to be a valid structured grids file!");
    std::string line, word;
    if (!std::getline(in, line) || line.compare(0, 5, "# vtk") != 0)
        throw std::logic_error("VTK file \"" + filename + "\" does not exist or does not "
                "seem to be a valid structured grids file!");
    std::getline(in, line);                    // title
    in >> word;
    if (word != "ASCII") throw std::logic_error("VTK file \"" + filename + "\": only ASCII legacy files are supported");

    long dims[3] = { 0, 0, 0 };
    double origin[3] = { 0, 0, 0 }, spacing[3] = { 1, 1, 1 };
    std::size_t n_points = 0;
    bool structured = false, have_data = false;
    while (in >> word) {
        if (word == "DATASET") {
            in >> word;
            structured = word == "STRUCTURED_POINTS";
        } else if (word == "DIMENSIONS") in >> dims[0] >> dims[1] >> dims[2];
        else if (word == "ORIGIN") in >> origin[0] >> origin[1] >> origin[2];
        else if (word == "SPACING" || word == "ASPECT_RATIO") in >> spacing[0] >> spacing[1] >> spacing[2];
        else if (word == "POINT_DATA") in >> n_points;
        else if (word == "SCALARS") std::getline(in, line);
        else if (word == "LOOKUP_TABLE") {
            in >> word;
            have_data = true;
            break;
        }
    }
    if (!structured)
        throw std::logic_error("VTK file \"" + filename + "\" does not exist or does not "
                "seem to be a valid structured grids file!");
    if (!have_data || dims[0] <= 0 || dims[1] <= 0 || dims[2] <= 0) throw std::logic_error("Point data is null!");
    const std::size_t expected = std::size_t(dims[0]) * dims[1] * dims[2];
    if (n_points != expected) throw std::logic_error("Could not read file!");

    auto domain = make_unique<Domain<lattice_model>>(dims[0], dims[1], dims[2], fluid_collision_model,
            origin[0], origin[1], origin[2], spacing[0], spacing[1], spacing[2]);
    auto& solid = BoundaryKeeper<lattice_model>::template get_collision<solid_collision_model>(*domain);
    // interior cells only, linearly, x fastest (io/vtk.hpp:141-150); the mask is painted on the device
    std::vector<std::uint8_t> mask(expected);
    for (std::size_t i = 0; i < expected; ++i) {
        int v;
        if (!(in >> v)) throw std::logic_error("Could not read file!");
        mask[i] = v != 0;
    }
    domain->apply_fluid_mask(mask.data(), solid);
    return domain;
}

} // namespace io
} // namespace lbm

// include/lbm/io/scenario.h -- scenario XML -> Domain with boundary conditions.
//
// Same entry point, element/attribute names, defaults, document-order application
// and error messages as the reference's include/io/scenario.h:21-188 (format
// description: build/scenarios/README.xml).  The reference parses with pugixml and
// boost::tokenizer/lexical_cast; neither is a dependency here -- the XML subset
// scenario files use (declaration, comments, nested elements, quoted attributes,
// self-closing tags) is read by the ~100-line parser below.
#pragma once
#include <cstdint>
#include <fstream>
#include <iostream>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "domain.h"
#include "vtk.h"
#include "boundary.h"
#include "configuration.h"

namespace lbm
{
namespace io
{

namespace xml
{

struct Node {
    std::string name;
    std::vector<std::pair<std::string, std::string>> attributes;
    std::vector<Node> children;

    const std::string* attribute(const std::string& key) const
    {
        for (const auto& a : attributes)
            if (a.first == key) return &a.second;
        return nullptr;
    }
    const Node* child(const std::string& tag) const
    {
        for (const auto& c : children)
            if (c.name == tag) return &c;
        return nullptr;
    }
};

class Parser
{
    const std::string& s;
    std::size_t p { 0 };

    [[noreturn]] void error(const std::string& what) const
    {
        throw std::logic_error("XML parse error near offset " + std::to_string(p) + ": " + what);
    }
    void skip_space()
    {
        while (p < s.size() && std::isspace(static_cast<unsigned char>(s[p]))) ++p;
    }
    bool starts(const char* lit) const { return s.compare(p, std::char_traits<char>::length(lit), lit) == 0; }
    void skip_misc()   // whitespace, comments, declarations, doctype
    {
        for (;;) {
            skip_space();
            if (starts("<!--")) {
                const auto e = s.find("-->", p + 4);
                if (e == std::string::npos) error("unterminated comment");
                p = e + 3;
            } else if (starts("<?")) {
                const auto e = s.find("?>", p + 2);
                if (e == std::string::npos) error("unterminated declaration");
                p = e + 2;
            } else if (starts("<!")) {
                const auto e = s.find('>', p);
                if (e == std::string::npos) error("unterminated markup");
                p = e + 1;
            } else {
                return;
            }
        }
    }
    std::string name()
    {
        const std::size_t b = p;
        while (p < s.size() && (std::isalnum(static_cast<unsigned char>(s[p])) || s[p] == '-' || s[p] == '_' || s[p] == ':' || s[p] == '.')) ++p;
        if (p == b) error("name expected");
        return s.substr(b, p - b);
    }
    static std::string unescape(const std::string& v)
    {
        std::string out;
        for (std::size_t i = 0; i < v.size(); ++i) {
            if (v[i] != '&') { out += v[i]; continue; }
            const auto e = v.find(';', i);
            const std::string ent = e == std::string::npos ? "" : v.substr(i + 1, e - i - 1);
            if (ent == "amp") out += '&';
            else if (ent == "lt") out += '<';
            else if (ent == "gt") out += '>';
            else if (ent == "quot") out += '"';
            else if (ent == "apos") out += '\'';
            else { out += v[i]; continue; }
            i = e;
        }
        return out;
    }

public:
    explicit Parser(const std::string& text) : s(text) {}

    Node element()
    {
        if (p >= s.size() || s[p] != '<') error("'<' expected");
        ++p;
        Node node;
        node.name = name();
        for (;;) {
            skip_space();
            if (p >= s.size()) error("unterminated tag");
            if (s[p] == '/') {
                if (p + 1 >= s.size() || s[p + 1] != '>') error("'/>' expected");
                p += 2;
                return node;
            }
            if (s[p] == '>') { ++p; break; }
            const std::string key = name();
            skip_space();
            if (p >= s.size() || s[p] != '=') error("'=' expected after attribute " + key);
            ++p;
            skip_space();
            if (p >= s.size() || (s[p] != '"' && s[p] != '\'')) error("quoted attribute value expected");
            const char quote = s[p++];
            const auto e = s.find(quote, p);
            if (e == std::string::npos) error("unterminated attribute value");
            node.attributes.emplace_back(key, unescape(s.substr(p, e - p)));
            p = e + 1;
        }
        for (;;) {   // content
            const auto lt = s.find('<', p);
            if (lt == std::string::npos) error("missing </" + node.name + ">");
            p = lt;
            if (starts("</")) {
                p += 2;
                const std::string closing = name();
                if (closing != node.name) error("</" + closing + "> closes <" + node.name + ">");
                skip_space();
                if (p >= s.size() || s[p] != '>') error("'>' expected");
                ++p;
                return node;
            }
            if (starts("<!--") || starts("<?") || starts("<!")) { skip_misc(); continue; }
            node.children.push_back(element());
        }
    }

    Node document()
    {
        skip_misc();
        Node root = element();
        skip_misc();
        return root;
    }
};

inline bool load_file(const std::string& filename, Node& root)
{
    std::ifstream in(filename, std::ios::binary);
    if (!in) return false;
    std::stringstream buffer;
    buffer << in.rdbuf();
    const std::string text = buffer.str();
    try {
        Node doc;
        doc.children.push_back(Parser(text).document());
        root = doc;
    } catch (const std::logic_error&) {
        return false;
    }
    return true;
}

inline double as_double(const std::string& v) { return std::strtod(v.c_str(), nullptr); }   // pugixml as_double
inline unsigned as_uint(const std::string& v) { return unsigned(std::strtoul(v.c_str(), nullptr, 10)); }

} // namespace xml

inline void check_attribute(const xml::Node& node, const std::string& name)
{
    if (!node.attribute(name))
        throw std::logic_error("Missing attribute \"" + name + "\" for node \"" + node.name + "\"!");
}

// one handler object per <boundary> node, owned by BoundaryKeeper (io/scenario.h:28-88)
template <typename lattice_model>
auto parse_condition(const xml::Node& boundary, Domain<lattice_model>& domain) -> NonFluidCollision<lattice_model>&
{
    using Keeper = BoundaryKeeper<lattice_model>;
    check_attribute(boundary, "condition");
    const std::string condition = *boundary.attribute("condition");

    auto velocity = [&]() {
        for (const char* key : { "vx", "vy", "vz" }) check_attribute(boundary, key);
        return double_array<lattice_model::D> { xml::as_double(*boundary.attribute("vx")),
            xml::as_double(*boundary.attribute("vy")), xml::as_double(*boundary.attribute("vz")) };
    };
    auto reference_density = [&]() {   // optional, default 1.0
        const std::string* rho = boundary.attribute("rho-ref");
        return rho ? xml::as_double(*rho) : 1.0;
    };

    if (condition == "noslip") return Keeper::template get_collision<NoSlipBoundary<lattice_model>>(domain);
    if (condition == "movingwall") {
        const auto wall_velocity = velocity();
        return Keeper::template get_collision<MovingWallBoundary<lattice_model>>(domain, wall_velocity);
    }
    if (condition == "freeslip") return Keeper::template get_collision<FreeSlipBoundary<lattice_model>>(domain);
    if (condition == "outflow") {
        const double rho_ref = reference_density();
        return Keeper::template get_collision<OutflowBoundary<lattice_model>>(domain, rho_ref);
    }
    if (condition == "inflow") {
        const auto inflow_velocity = velocity();
        const double rho_ref = reference_density();
        return Keeper::template get_collision<InflowBoundary<lattice_model>>(domain, inflow_velocity, rho_ref);
    }
    if (condition == "pressure") {
        check_attribute(boundary, "rho-in");
        const double rho_in = xml::as_double(*boundary.attribute("rho-in"));
        return Keeper::template get_collision<PressureBoundary<lattice_model>>(domain, rho_in);
    }
    if (condition == "periodic")   // extension, see PeriodicBoundary in boundary.h
        return Keeper::template get_collision<PeriodicBoundary<lattice_model>>(domain);
    throw std::logic_error(condition + " boundary condition not supported!");
}

// extents: x0/xmax/y0/ymax/z0/zmax = whole ghost planes (edges included), or six inclusive
// indices "xbegin xend ybegin yend zbegin zend" (io/scenario.h:91-128)
template <typename lattice_model>
void parse_boundary(const xml::Node& boundary, Domain<lattice_model>& domain)
{
    check_attribute(boundary, "extent");
    const std::string extent = *boundary.attribute("extent");
    auto& condition = parse_condition<lattice_model>(boundary, domain);
    const auto xl = domain.xlength(), yl = domain.ylength(), zl = domain.zlength();
    static const std::map<std::string, int> faces = { { "x0", 0 }, { "xmax", 1 }, { "y0", 2 }, { "ymax", 3 }, { "z0", 4 }, { "zmax", 5 } };
    const auto face = faces.find(extent);
    if (face != faces.end()) {
        std::size_t lo[3] = { 0, 0, 0 }, hi[3] = { xl + 1, yl + 1, zl + 1 };
        const int axis = face->second / 2;
        if (face->second % 2) lo[axis] = hi[axis];
        else hi[axis] = 0;
        domain.setBoundaryCondition(condition, lo[0], hi[0], lo[1], hi[1], lo[2], hi[2]);
        return;
    }
    std::vector<std::uint64_t> ex;
    std::istringstream tokens(extent);
    std::string token;
    while (tokens >> token) {
        std::size_t used = 0;
        unsigned long long v = 0;
        try {
            v = std::stoull(token, &used);
        } catch (const std::exception&) {
            used = 0;
        }
        if (used != token.size() || token[0] == '-')
            throw std::logic_error("bad lexical cast: source type value could not be interpreted as target");
        ex.push_back(v);
    }
    if (ex.size() < 6)
        throw std::logic_error("Extent \"" + extent + "\" is not complete! Must be six values!");
    domain.setBoundaryCondition(condition, ex[0], ex[1], ex[2], ex[3], ex[4], ex[5]);
}

template <typename lattice_model>
auto parse_scenario_file(const std::string& filename, Config& cfg, FluidCollision<lattice_model>& collision)
    -> std::unique_ptr<Domain<lattice_model>>
{
    xml::Node doc;
    if (!xml::load_file(filename, doc))
        throw std::logic_error("XML file \"" + filename + "\" could not be read properly!");
    const xml::Node* scenario = doc.child("scenario");
    if (!scenario) throw std::logic_error("Scenario node missing!");
    if (!scenario->attribute("name")) throw std::logic_error("Scenario name is missing!");
    const std::string scenario_name = *scenario->attribute("name");
    std::cout << "Reading scenario: \"" << scenario_name << "\"..." << std::endl;
    cfg.set_output_filename(scenario_name);

    const xml::Node* xml_domain = scenario->child("domain");
    if (!xml_domain) throw std::logic_error("Domain node is missing!");
    const std::string* vtk_file = xml_domain->attribute("vtk-file");
    const std::string* xml_xl = xml_domain->attribute("xl");
    const std::string* xml_yl = xml_domain->attribute("yl");
    const std::string* xml_zl = xml_domain->attribute("zl");

    std::unique_ptr<Domain<lattice_model>> domain;
    if (vtk_file) {
        domain = read_vtk_point_file<lattice_model, NoSlipBoundary<lattice_model>>(*vtk_file, collision);
    } else if (xml_xl && xml_yl && xml_zl) {
        const unsigned xl = xml::as_uint(*xml_xl), yl = xml::as_uint(*xml_yl), zl = xml::as_uint(*xml_zl);
        if (!(xl > 0 && yl > 0 && zl > 0)) throw std::logic_error("Domain lengths xl/yl/zl must be positive!");
        domain = make_unique<Domain<lattice_model>>(xl, yl, zl, collision);
    } else {
        throw std::logic_error("Neither vtk-file nor xl/yl/zl attribute provided to domain node!");
    }
    for (const auto& node : xml_domain->children)
        if (node.name == "boundary") parse_boundary(node, *domain);
    return domain;
}

} // namespace io
} // namespace lbm

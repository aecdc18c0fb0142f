// include/lbm/domain.h -- Domain<M>: the lattice, resident on one or more B200s.
//
// Public interface of the reference's include/domain.h:22-55 (constructor,
// getters, idx/in_bounds/cell, setBoundaryCondition, set_nonfluid_cells_nullcollide,
// stream/collide/swap, create_subdomain).  What changes underneath:
//   * storage: the two Lattice_fields of domain.h:13-14 become device-resident
//     padded structure-of-arrays buffers owned through the C ABI (lbm_b200.h);
//   * stream(); swap(); collide();  (src/main.cpp:50-52) is ONE fused kernel launch,
//     issued by collide() once the three calls have been made in that order;
//   * the lattice is split into z-slabs over `lbm::device::gpus()` GPUs; the sweep
//     stores the populations leaving a slab into the neighbour's ghost plane over
//     NVLink (what parallel.h / create_subdomain only sketched);
//   * Domain::cell() serves a host mirror that is downloaded on demand and uploaded
//     before the next step if a mutable reference was handed out -- set-up,
//     inspection and tests only, never inside the time loop.
// Errors of the C ABI surface as std::runtime_error, so main's catch block
// (src/main.cpp:68-71) behaves as before.
#pragma once
#include <cassert>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "lbmdefinitions.h"
#include "collision.h"
#include "parallel.h"
#include "helper.h"

namespace lbm
{

namespace device
{
// process-wide knobs the fixed Domain constructor signature has no room for
inline int& gpus_ref()
{
    static int n = [] { const char* e = std::getenv("LBM_B200_GPUS"); return e ? std::max(1, std::atoi(e)) : 1; }();
    return n;
}
inline int& arithmetic_ref()
{
    static int m = [] {
        const char* e = std::getenv("LBM_B200_ARITHMETIC");
        return (e && std::string(e) == "exact") ? LBM_B200_EXACT : LBM_B200_FAST;
    }();
    return m;
}
// Handlers set through Domain::cell() -- the VTK mask reader (io/vtk.hpp:145-146),
// set_nonfluid_cells_nullcollide (domain.hpp:108-109), Cell::set_collision_handler -- reach only the
// collide field in the reference, so such a cell is solid on every other step.  literal = true
// reproduces that bit for bit; the default applies them to both lattices (DESIGN.md "deviations").
inline bool& mask_literal_ref()
{
    static bool b = [] { const char* e = std::getenv("LBM_B200_MASK_LITERAL"); return e && std::atoi(e) != 0; }();
    return b;
}
inline void set_gpus(int n) { gpus_ref() = n < 1 ? 1 : n; }
inline int gpus() { return gpus_ref(); }
inline void set_arithmetic(int mode) { arithmetic_ref() = mode; }
inline int arithmetic() { return arithmetic_ref(); }
inline void set_mask_literal(bool on) { mask_literal_ref() = on; }
inline bool mask_literal() { return mask_literal_ref(); }

inline void check(int rc, const char* what)
{
    if (rc != 0) throw std::runtime_error(std::string(what) + ": " + lbm_b200_last_error());
}
} // namespace device

template <typename lattice_model>
class Domain
{
    const FluidCollision<lattice_model>* const collision;
    const std::size_t xl { 0 }, yl { 0 }, zl { 0 };
    double xo, yo, zo;
    double xs, ys, zs;

    struct Slab {
        lbm_b200_t* handle;
        std::size_t z_first, zl_local;      // first plane and number of planes along the split axis
    };
    std::vector<Slab> slabs;
    // Split axis of a multi-GPU domain: z-slabs (x-y planes) normally; y-slabs (x-z planes) when the z extent is
    // shorter than the number of GPUs (LBM_B200_SPLIT_AXIS=y|z overrides).  A y-slab's host arrays cover its own
    // rows of every x-y plane, so the plane transfers below are stitched row-wise.
    int split_axis { LBM_B200_AXIS_Z };
    bool y_split() const { return split_axis == LBM_B200_AXIS_Y && slabs.size() > 1; }

    // one x-y plane (ghost shell included) of per-cell items of `width` values: global <-> the rows a y-slab stores
    template <typename T>
    void rows_to_global(const Slab& s, const T* local, T* global, std::size_t width) const
    {
        const std::size_t row = (xl + 2) * width;
        const bool first = s.z_first == 1, last = s.z_first + s.zl_local - 1 == yl;
        for (std::size_t ly = first ? 0 : 1; ly <= s.zl_local + (last ? 1 : 0); ++ly)
            std::memcpy(global + (s.z_first - 1 + ly) * row, local + ly * row, row * sizeof(T));
    }
    template <typename T>
    void rows_to_local(const Slab& s, const T* global, T* local, std::size_t width) const
    {
        const std::size_t row = (xl + 2) * width;
        std::memcpy(local, global + (s.z_first - 1) * row, (s.zl_local + 2) * row * sizeof(T));
    }
    std::size_t slab_plane_cells(const Slab& s) const { return (xl + 2) * (s.zl_local + 2); }

    // Handler bookkeeping.  The per-cell handler maps live on the device (one kind byte and one handler id
    // per cell and lattice); the host keeps the table id -> handler object (id 0 is the fluid operator)
    // and a journal of boxes not yet sent.  Applying a scenario to a 512^3 lattice therefore ships a few
    // hundred bytes; per-cell questions (cell().get_collision_handler(), handler()) fetch one x-y plane
    // of the maps on demand.
    std::vector<const Collision<lattice_model>*> handlers;
    std::size_t handlers_pushed { 0 };
    std::vector<std::uint64_t> pending_boxes;       // 6 inclusive indices per box
    std::vector<std::uint16_t> pending_ids;
    struct HandlerPlane {
        std::vector<std::uint8_t> kind;
        std::vector<std::uint16_t> id;
        bool valid { false };
    };
    mutable std::vector<HandlerPlane> handler_planes;

    // Host mirror behind cell(): one x-y plane at a time, downloaded on first access after a step
    // and uploaded before the next step if a mutable reference into it was handed out.  Touching
    // a few cells of a 512^3 lattice therefore moves a few 42 MB planes, not the 20 GB lattice.
    // Plane storage is kept across steps, so references stay valid (and are refreshed when the
    // plane is fetched again), like references into the reference's Cell vectors.
    struct PlaneMirror {
        Lattice_field<lattice_model> cells;
        std::vector<const Collision<lattice_model>*> fetched;   // handlers as reported when the plane was fetched
        bool valid { false };
        bool dirty { false };
    };
    mutable std::vector<PlaneMirror> planes;
    mutable bool readback_ready { false };   // edge planes pushed across slab cuts for this time level

    bool streamed { false }, swapped { false };
    std::uint64_t steps_done { 0 };

    static NullCollision<lattice_model>& null_collision()
    {
        static NullCollision<lattice_model> instance;      // domain.hpp:103
        return instance;
    }

    std::size_t plane_cells() const { return (xl + 2) * (yl + 2); }

    std::uint16_t intern(const Collision<lattice_model>* h)
    {
        for (std::size_t i = 0; i < handlers.size(); ++i)
            if (handlers[i] == h) return std::uint16_t(i);
        if (h->device_kind() < 0)
            throw std::logic_error("this collision operator only exists as host code (device_kind() < 0); "
                    "user-defined operators cannot run in the CUDA sweep and there is no CPU fallback");
        if (handlers.size() >= 65535) throw std::logic_error("more than 65535 distinct collision handlers");
        handlers.push_back(h);
        return std::uint16_t(handlers.size() - 1);
    }

    // handler object for what the device reports about a cell
    const Collision<lattice_model>* decode(std::uint8_t kind, std::uint16_t id) const
    {
        const Collision<lattice_model>* h = id < handlers.size() ? handlers[id] : collision;
        // a cell tagged by set_nonfluid_cells_nullcollide keeps its former handler's id
        if (kind == LBM_B200_NULL && h->device_kind() != LBM_B200_NULL) return &null_collision();
        return h;
    }

    void push_handlers()
    {
        if (handlers_pushed == handlers.size()) return;
        std::vector<lbm_b200_bc> table(handlers.size());
        for (std::size_t i = 0; i < handlers.size(); ++i) table[i] = handlers[i]->descriptor();
        for (auto& s : slabs) device::check(lbm_b200_set_handlers(s.handle, table.data(), int(table.size())), "lbm_b200_set_handlers");
        handlers_pushed = handlers.size();
    }

    // journal -> device: the handler table and the boxes of setBoundaryCondition, in order
    void push_geometry()
    {
        push_handlers();
        if (pending_ids.empty()) return;
        for (auto& s : slabs)
            device::check(lbm_b200_paint_boxes(s.handle, pending_boxes.data(), pending_ids.data(), int(pending_ids.size())),
                    "lbm_b200_paint_boxes");
        pending_boxes.clear();
        pending_ids.clear();
    }

    void invalidate_handler_planes() const
    {
        for (auto& hp : handler_planes) hp.valid = false;
    }

    // boundary cells next to a slab cut need the neighbour slab's full edge plane for the read-back pass
    void prepare_readback() const
    {
        if (slabs.size() < 2) return;
        for (auto& s : slabs) device::check(lbm_b200_sync(s.handle), "lbm_b200_sync");
        for (auto& s : slabs) device::check(lbm_b200_halo_push_all(s.handle), "lbm_b200_halo_push_all");
        for (auto& s : slabs) device::check(lbm_b200_sync(s.handle), "lbm_b200_sync");
        for (auto& s : slabs) device::check(lbm_b200_halo_pushed(s.handle), "lbm_b200_halo_pushed");
    }

    // slab that owns global plane z (the two physical ghost planes belong to the first / last slab)
    std::size_t owner_of(std::size_t z) const
    {
        for (std::size_t i = 0; i < slabs.size(); ++i)
            if (z >= slabs[i].z_first && z < slabs[i].z_first + slabs[i].zl_local) return i;
        return z == 0 ? 0 : slabs.size() - 1;
    }

    // handler maps of plane z as Domain::cell() of the reference would report them
    const HandlerPlane& fetch_handlers(std::size_t z) const
    {
        HandlerPlane& hp = handler_planes[z];
        if (hp.valid) return hp;
        const_cast<Domain*>(this)->push_geometry();
        hp.kind.resize(plane_cells());
        hp.id.resize(plane_cells());
        if (y_split()) {
            std::vector<std::uint8_t> k;
            std::vector<std::uint16_t> id;
            for (const Slab& s : slabs) {
                k.resize(slab_plane_cells(s));
                id.resize(slab_plane_cells(s));
                device::check(lbm_b200_get_geometry_planes(s.handle, k.data(), id.data(), z, 1), "lbm_b200_get_geometry_planes");
                rows_to_global(s, k.data(), hp.kind.data(), 1);
                rows_to_global(s, id.data(), hp.id.data(), 1);
            }
        } else {
            const Slab& s = slabs[owner_of(z)];
            device::check(lbm_b200_get_geometry_planes(s.handle, hp.kind.data(), hp.id.data(), z - (s.z_first - 1), 1),
                    "lbm_b200_get_geometry_planes");
        }
        hp.valid = true;
        return hp;
    }

    // mirror <- device
    void fetch_plane(std::size_t z) const
    {
        PlaneMirror& pm = planes[z];
        if (pm.valid) return;
        auto* self = const_cast<Domain*>(this);
        self->push_geometry();
        if (!readback_ready) {
            prepare_readback();
            readback_ready = true;
        }
        constexpr std::size_t Q = lattice_model::Q;
        if (pm.cells.empty()) pm.cells.assign(plane_cells(), Cell<lattice_model>(collision));
        pm.fetched.resize(plane_cells());
        std::vector<double> buf(plane_cells() * Q);
        if (y_split()) {
            std::vector<double> part;
            for (const Slab& s : slabs) {
                part.resize(slab_plane_cells(s) * Q);
                device::check(lbm_b200_download_planes(s.handle, part.data(), LBM_B200_COLLIDE_FIELD, z, 1), "lbm_b200_download_planes");
                rows_to_global(s, part.data(), buf.data(), Q);
            }
        } else {
            const Slab& s = slabs[owner_of(z)];
            device::check(lbm_b200_download_planes(s.handle, buf.data(), LBM_B200_COLLIDE_FIELD, z - (s.z_first - 1), 1),
                    "lbm_b200_download_planes");
        }
        const HandlerPlane& hp = fetch_handlers(z);
        for (std::size_t c = 0; c < plane_cells(); ++c) {
            std::memcpy(pm.cells[c].data(), buf.data() + c * Q, Q * sizeof(double));
            pm.fetched[c] = decode(hp.kind[c], hp.id[c]);
            pm.cells[c].set_collision_handler(pm.fetched[c]);
        }
        pm.valid = true;
        pm.dirty = false;
    }

    // mirror -> device: planes that may have been written, to every slab that stores them
    void push_mirror()
    {
        constexpr std::size_t Q = lattice_model::Q;
        std::vector<double> buf;
        for (std::size_t z = 0; z < planes.size(); ++z) {
            PlaneMirror& pm = planes[z];
            if (!pm.dirty) continue;
            // handlers changed through Cell::set_collision_handler
            std::vector<std::size_t> changed;
            for (std::size_t c = 0; c < plane_cells(); ++c)
                if (pm.cells[c].get_collision_handler() != pm.fetched[c]) changed.push_back(c);
            if (!changed.empty()) {
                push_geometry();                       // earlier boxes first
                if (device::mask_literal()) {
                    // the reference writes the collide field's handler only (cell.h:62-63 through domain.hpp:74-83)
                    HandlerPlane hp = fetch_handlers(z);
                    for (auto c : changed) {
                        const auto* h = pm.cells[c].get_collision_handler();
                        hp.id[c] = intern(h);
                        hp.kind[c] = std::uint8_t(h->device_kind());
                    }
                    push_handlers();
                    for (auto& s : slabs)
                        if (!y_split() && z + 1 >= s.z_first && z <= s.z_first + s.zl_local)
                            device::check(lbm_b200_set_geometry_planes(s.handle, hp.kind.data(), hp.id.data(), z - (s.z_first - 1), 1, 1),
                                    "lbm_b200_set_geometry_planes");
                    if (y_split()) throw std::logic_error("literal handler edits are limited to single-GPU domains");
                } else {
                    for (auto c : changed) {
                        const std::uint64_t x = c % (xl + 2), y = c / (xl + 2);
                        const std::uint64_t box[6] = { x, x, y, y, z, z };
                        pending_boxes.insert(pending_boxes.end(), box, box + 6);
                        pending_ids.push_back(intern(pm.cells[c].get_collision_handler()));
                    }
                    push_geometry();
                }
                for (auto c : changed) pm.fetched[c] = pm.cells[c].get_collision_handler();
                handler_planes[z].valid = false;
            }
            buf.resize(plane_cells() * Q);
            for (std::size_t c = 0; c < plane_cells(); ++c) std::memcpy(buf.data() + c * Q, pm.cells[c].data(), Q * sizeof(double));
            if (y_split()) {
                std::vector<double> part;
                for (auto& s : slabs) {                                   // every y-slab stores its rows of this plane
                    part.resize(slab_plane_cells(s) * Q);
                    rows_to_local(s, buf.data(), part.data(), Q);
                    device::check(lbm_b200_upload_planes(s.handle, part.data(), LBM_B200_COLLIDE_FIELD, z, 1), "lbm_b200_upload_planes");
                }
            } else {
                for (auto& s : slabs)
                    if (z + 1 >= s.z_first && z <= s.z_first + s.zl_local)   // own plane or one of its two ghost planes
                        device::check(lbm_b200_upload_planes(s.handle, buf.data(), LBM_B200_COLLIDE_FIELD, z - (s.z_first - 1), 1),
                                "lbm_b200_upload_planes");
            }
            pm.dirty = false;
        }
    }

    void invalidate_mirror()
    {
        for (auto& pm : planes) pm.valid = false;
        readback_ready = false;
    }

    void require_idle(const char* what) const
    {
        if (streamed || swapped)
            throw std::logic_error(std::string(what) + " between stream() and collide(): the three calls "
                    "stream(); swap(); collide(); execute as one fused device step");
    }

public:
    Domain(std::size_t xl, std::size_t yl, std::size_t zl, FluidCollision<lattice_model>& _collision,
            double xorigin = 0, double yorigin = 0, double zorigin = 0,
            double xspacing = 1, double yspacing = 1, double zspacing = 1)
        : collision { &_collision }, xl { xl }, yl { yl }, zl { zl },
          xo { xorigin }, yo { yorigin }, zo { zorigin }, xs { xspacing }, ys { yspacing }, zs { zspacing }
    {
        const auto* bgk = dynamic_cast<const BGKCollision<lattice_model>*>(collision);
        if (!bgk)
            throw std::logic_error("Domain: only BGKCollision can run in the CUDA sweep (the reference asserts "
                    "collision-model == bgk as well, io/configuration.h:126-128)");
        handlers.push_back(collision);
        handlers_pushed = 1;            // a fresh handle's table already holds the fluid operator as id 0
        planes.resize(zl + 2);
        handler_planes.resize(zl + 2);
        int n = device::gpus();
        if (const char* e = std::getenv("LBM_B200_SPLIT_AXIS")) split_axis = (e[0] == 'y' || e[0] == 'Y') ? LBM_B200_AXIS_Y : LBM_B200_AXIS_Z;
        else if (std::size_t(n) > zl && yl >= zl) split_axis = LBM_B200_AXIS_Y;      // too few x-y planes: cut along y
        const std::size_t split_len = split_axis == LBM_B200_AXIS_Y ? yl : zl;
        if (std::size_t(n) > split_len) n = int(split_len);
        const int visible = lbm_b200_device_count();
        if (visible < 1) throw std::runtime_error("Domain: no CUDA device visible; there is no CPU fallback");
        if (n > visible) throw std::runtime_error("Domain: " + std::to_string(n) + " GPUs requested, " + std::to_string(visible) + " visible");
        std::size_t z = 1;
        for (int r = 0; r < n; ++r) {
            const std::size_t nz = split_len / n + (std::size_t(r) < split_len % n ? 1 : 0);
            lbm_b200_t* h = nullptr;
            const int rc = lbm_b200_create_slab_axis(&h, int(lattice_model::Q), xl, yl, zl, n > 1 ? split_axis : LBM_B200_AXIS_Z,
                    z, nz, bgk->relaxation_time(), n > 1 ? r : -1);
            if (rc != 0) {
                const std::string msg = lbm_b200_last_error();
                for (auto& s : slabs) lbm_b200_destroy(s.handle);
                throw std::runtime_error("lbm_b200_create_slab: " + msg);
            }
            slabs.push_back({ h, z, nz });
            device::check(lbm_b200_set_arithmetic(h, device::arithmetic()), "lbm_b200_set_arithmetic");
            z += nz;
        }
        for (std::size_t r = 0; r + 1 < slabs.size(); ++r) {
            device::check(lbm_b200_connect_local(slabs[r].handle, LBM_B200_UP, slabs[r + 1].handle), "lbm_b200_connect_local");
            device::check(lbm_b200_connect_local(slabs[r + 1].handle, LBM_B200_DOWN, slabs[r].handle), "lbm_b200_connect_local");
        }
    }

    ~Domain()
    {
        for (auto& s : slabs) lbm_b200_sync(s.handle);
        for (auto& s : slabs) lbm_b200_disconnect(s.handle);
        for (auto& s : slabs) lbm_b200_destroy(s.handle);
    }
    Domain(const Domain&) = delete;
    Domain& operator=(const Domain&) = delete;

    // Getters
    auto xlength() const -> decltype(xl) { return xl; }
    auto ylength() const -> decltype(yl) { return yl; }
    auto zlength() const -> decltype(zl) { return zl; }
    auto xorigin() const -> decltype(xo) { return xo; }
    auto yorigin() const -> decltype(yo) { return yo; }
    auto zorigin() const -> decltype(zo) { return zo; }
    auto xspacing() const -> decltype(xs) { return xs; }
    auto yspacing() const -> decltype(ys) { return ys; }
    auto zspacing() const -> decltype(zs) { return zs; }

    // Helper functions
    auto idx(int x, int y, int z) const -> int { return x + int(xl + 2) * y + int((xl + 2) * (yl + 2)) * z; }
    auto in_bounds(int x, int y, int z) const -> bool
    {
        return x > 0 && x < int(xl) + 1 && y > 0 && y < int(yl) + 1 && z > 0 && z < int(zl) + 1;
    }
    auto cell(int x, int y, int z) const -> const Cell<lattice_model>&
    {
        require_idle("Domain::cell()");
        fetch_plane(std::size_t(z));
        return planes[std::size_t(z)].cells[std::size_t(x) + (xl + 2) * std::size_t(y)];
    }
    auto cell(int x, int y, int z) -> Cell<lattice_model>&
    {
        require_idle("Domain::cell()");
        fetch_plane(std::size_t(z));
        planes[std::size_t(z)].dirty = true;   // a mutable reference escapes: assume it is written
        return planes[std::size_t(z)].cells[std::size_t(x) + (xl + 2) * std::size_t(y)];
    }

    // handler of a cell without touching the population mirror
    auto handler(int x, int y, int z) const -> const Collision<lattice_model>*
    {
        const std::size_t c = std::size_t(x) + (xl + 2) * std::size_t(y);
        const PlaneMirror& pm = planes[std::size_t(z)];
        if (pm.valid && pm.dirty) return pm.cells[c].get_collision_handler();
        const HandlerPlane& hp = fetch_handlers(std::size_t(z));
        return decode(hp.kind[c], hp.id[c]);
    }

    // domain.hpp:101-113: interior non-fluid cells without any interior fluid neighbour get the
    // do-nothing handler.  A pure optimisation in the reference; the device sweep never visits
    // such cells anyway, so only the handler maps change (on the device, one kernel).
    auto set_nonfluid_cells_nullcollide() -> void
    {
        require_idle("set_nonfluid_cells_nullcollide()");
        push_mirror();
        push_geometry();
        for (auto& s : slabs) {
            std::uint64_t tagged = 0;
            device::check(lbm_b200_tag_null_cells(s.handle, device::mask_literal() ? 1 : 0, &tagged), "lbm_b200_tag_null_cells");
        }
        invalidate_handler_planes();
        for (auto& pm : planes) pm.valid = false;     // handlers of mirrored cells are refreshed on the next access
    }

    // domain.hpp:175-194: inclusive box, later calls overwrite earlier ones
    auto setBoundaryCondition(NonFluidCollision<lattice_model>& condition,
            std::size_t x0, std::size_t xE, std::size_t y0, std::size_t yE, std::size_t z0, std::size_t zE) -> void
    {
        require_idle("setBoundaryCondition()");
        if (!(xE >= x0 && yE >= y0 && zE >= z0) || !(xE < xl + 2 && yE < yl + 2 && zE < zl + 2))
            throw std::out_of_range("setBoundaryCondition: extent outside the domain (the reference asserts this, domain.hpp:180-181)");
        push_mirror();
        const std::uint16_t id = intern(&condition);
        const std::uint64_t box[6] = { x0, xE, y0, yE, z0, zE };
        pending_boxes.insert(pending_boxes.end(), box, box + 6);
        pending_ids.push_back(id);
        for (auto z = z0; z <= zE; ++z) {
            handler_planes[z].valid = false;
            PlaneMirror& pm = planes[z];
            if (!pm.valid) continue;
            for (auto y = y0; y <= yE; ++y)
                for (auto x = x0; x <= xE; ++x) {
                    pm.cells[x + (xl + 2) * y].set_collision_handler(&condition);
                    pm.fetched[x + (xl + 2) * y] = &condition;
                }
        }
    }

    // the mask loop of io/vtk.hpp:141-150 in one call: interior cells whose mask byte is 0 (x fastest, then y,
    // then z, like the POINT_DATA of the file) take `solid`.  With device::mask_literal() the handler reaches
    // the collide field only, as Domain::cell(x,y,z).set_collision_handler(&solid) does in the reference.
    auto apply_fluid_mask(const std::uint8_t* mask, NonFluidCollision<lattice_model>& solid) -> void
    {
        require_idle("apply_fluid_mask()");
        push_mirror();
        const std::uint16_t id = intern(&solid);
        push_geometry();
        for (auto& s : slabs)
            device::check(lbm_b200_paint_mask(s.handle, mask, id, device::mask_literal() ? 1 : 0), "lbm_b200_paint_mask");
        invalidate_handler_planes();
        for (auto& pm : planes) pm.valid = false;
    }

    // Iteration functions.  The reference runs three host loops (domain.hpp:116-172); here
    // stream() and swap() only record that they were called and collide() launches the fused
    // pull-stream + BGK + boundary sweep for one time step.
    auto stream() -> void
    {
        if (streamed) throw std::logic_error("stream() called twice without collide()");
        streamed = true;
    }
    auto swap() -> void
    {
        if (!streamed || swapped) throw std::logic_error("swap() must follow stream() (src/main.cpp:50-52 order)");
        swapped = true;
    }
    auto collide() -> void
    {
        if (!streamed || !swapped)
            throw std::logic_error("collide() must follow stream(); swap(); -- a stand-alone collide has no "
                    "device implementation");
        streamed = swapped = false;
        step(1);
    }

    // --- extensions ---
    // n fused time steps; asynchronous, later read-outs synchronise
    auto step(std::uint64_t n) -> void
    {
        require_idle("step()");
        push_mirror();
        push_geometry();
        if (slabs.size() == 1) {
            device::check(lbm_b200_step(slabs[0].handle, n), "lbm_b200_step");
        } else {
            std::vector<lbm_b200_t*> hs;
            for (auto& s : slabs) hs.push_back(s.handle);
            device::check(lbm_b200_step_group(hs.data(), int(hs.size()), n), "lbm_b200_step_group");
        }
        steps_done += n;
        if (n > 0) {
            invalidate_mirror();
            invalidate_handler_planes();     // tags of set_nonfluid_cells_nullcollide alternate with the swaps
        }
    }
    auto synchronize() const -> void
    {
        for (auto& s : slabs) device::check(lbm_b200_sync(s.handle), "lbm_b200_sync");
    }
    // density / velocity of the interior cells in z,y,x order (the loop of io/vtk.hpp:62-73),
    // reduced on the device and copied back
    auto macroscopic(double* rho, double* u) const -> void
    {
        require_idle("macroscopic()");
        auto* self = const_cast<Domain*>(this);
        self->push_mirror();
        self->push_geometry();
        if (!readback_ready) {
            prepare_readback();
            readback_ready = true;
        }
        if (y_split()) {
            // a y-slab returns its rows of every plane: gather through a temporary and place them row-wise
            std::vector<double> r, v;
            for (auto& s : slabs) {
                const std::size_t n = xl * s.zl_local * zl;
                if (rho) r.resize(n);
                if (u) v.resize(3 * n);
                device::check(lbm_b200_macroscopic(s.handle, rho ? r.data() : nullptr, u ? v.data() : nullptr), "lbm_b200_macroscopic");
                for (std::size_t z = 0; z < zl; ++z) {
                    const std::size_t dst = (z * yl + (s.z_first - 1)) * xl, src = z * s.zl_local * xl, cnt = s.zl_local * xl;
                    if (rho) std::memcpy(rho + dst, r.data() + src, cnt * sizeof(double));
                    if (u) std::memcpy(u + 3 * dst, v.data() + 3 * src, 3 * cnt * sizeof(double));
                }
            }
            return;
        }
        for (auto& s : slabs) {
            const std::size_t off = (s.z_first - 1) * xl * yl;
            device::check(lbm_b200_macroscopic(s.handle, rho ? rho + off : nullptr, u ? u + 3 * off : nullptr),
                    "lbm_b200_macroscopic");
        }
    }
    // split form: reduce now, copy while later steps run, wait in macroscopic_end() (rho / u should be
    // page-locked, see lbm_b200_host_alloc).  Multi-slab domains: call synchronize-free after macroscopic()
    // has been used once for this time level only if no boundary cell next to a cut matters.
    auto macroscopic_begin(double* rho, double* u) const -> void
    {
        require_idle("macroscopic_begin()");
        auto* self = const_cast<Domain*>(this);
        self->push_mirror();
        self->push_geometry();
        if (!readback_ready) {
            prepare_readback();
            readback_ready = true;
        }
        if (y_split()) {                 // rows are interleaved in the output: no asynchronous form, read out now
            macroscopic(rho, u);
            return;
        }
        for (auto& s : slabs) {
            const std::size_t off = (s.z_first - 1) * xl * yl;
            device::check(lbm_b200_macroscopic_begin(s.handle, rho ? rho + off : nullptr, u ? u + 3 * off : nullptr),
                    "lbm_b200_macroscopic_begin");
        }
    }
    auto macroscopic_end() const -> void
    {
        if (y_split()) return;
        for (auto& s : slabs) device::check(lbm_b200_macroscopic_end(s.handle), "lbm_b200_macroscopic_end");
    }
    auto split_axis_name() const -> const char* { return y_split() ? "y" : "z"; }
    auto gpu_count() const -> std::size_t { return slabs.size(); }
    auto timesteps_done() const -> std::uint64_t { return steps_done; }
    auto slab_handle(std::size_t i) const -> lbm_b200_t* { return slabs.at(i).handle; }

    // Parallelization tools.  The reference's x-slab helper (domain.hpp:197-248) copies an
    // x-interval into a new Domain and tags the cut faces with ParallelBoundary; nothing ever
    // exchanges data across them.  Kept for interface compatibility (host-side copy through
    // cell()); real multi-GPU runs use the built-in z-slab split instead.
    auto create_subdomain(parallel::ParallelBoundary<lattice_model>& parallel_boundary,
            std::size_t xstart, std::size_t xend, int rank, int number_of_ranks) const -> Domain_ptr<lattice_model>
    {
        if (!(xend > xstart)) throw std::out_of_range("create_subdomain: xend must exceed xstart");
        const std::size_t new_xl = xend - xstart + 1;
        auto sub = make_unique<Domain<lattice_model>>(new_xl, yl, zl,
                const_cast<FluidCollision<lattice_model>&>(*collision), xo + rank * new_xl * xs, yo, zo, xs, ys, zs);
        for (std::size_t z = 0; z < zl + 2; ++z)
            for (std::size_t y = 0; y < yl + 2; ++y)
                for (std::size_t x = xstart; x < xend; ++x)
                    sub->cell(int(x - xstart + 1), int(y), int(z)) = cell(int(x), int(y), int(z));
        const bool leftmost = rank == 0, rightmost = rank == number_of_ranks - 1;
        for (std::size_t z = 0; z < zl + 2; ++z)
            for (std::size_t y = 0; y < yl + 2; ++y) {
                if (leftmost) sub->cell(0, int(y), int(z)) = cell(0, int(y), int(z));
                if (rightmost) sub->cell(int(new_xl + 1), int(y), int(z)) = cell(int(xl + 1), int(y), int(z));
            }
        sub->push_mirror();
        if (!leftmost) sub->setBoundaryCondition(parallel_boundary, 0, 0, 0, yl + 1, 0, zl + 1);
        if (!rightmost) sub->setBoundaryCondition(parallel_boundary, new_xl + 1, new_xl + 1, 0, yl + 1, 0, zl + 1);
        return sub;
    }
};

template <typename lattice_model>
auto Cell<lattice_model>::has_fluid_vicinity(const Domain<lattice_model>& domain,
        const uint_array<lattice_model::D>& position) const -> bool
{
    for (std::size_t q = 0; q < lattice_model::Q; ++q) {
        const int nx = int(position[0]) + int(lattice_model::velocities[q][0]);
        const int ny = int(position[1]) + int(lattice_model::velocities[q][1]);
        const int nz = int(position[2]) + int(lattice_model::velocities[q][2]);
        if (domain.in_bounds(nx, ny, nz) && domain.handler(nx, ny, nz)->is_fluid()) return true;
    }
    return false;
}

} // namespace lbm

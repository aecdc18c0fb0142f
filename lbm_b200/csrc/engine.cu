// engine.cu -- host side of the C ABI declared in include/lbm_b200.h.
//
// One handle = one Domain (domain.h:12-19) or one z-slab of it, resident on one
// B200.  The reference's stream(); swap(); collide(); (src/main.cpp:50-52) is a
// single launch of sweep_kernel per step plus a buffer-index flip.
#include <cuda_runtime.h>
#include <sched.h>
#include <unistd.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/lbm_b200.h"
#include "kernels.cuh"

using namespace lbmb200;

// the device code's kind numbering is the ABI's
static_assert(K_FLUID == LBM_B200_FLUID && K_NOSLIP == LBM_B200_NOSLIP && K_MOVINGWALL == LBM_B200_MOVINGWALL
              && K_FREESLIP == LBM_B200_FREESLIP && K_OUTFLOW == LBM_B200_OUTFLOW && K_INFLOW == LBM_B200_INFLOW
              && K_PRESSURE == LBM_B200_PRESSURE && K_NULL == LBM_B200_NULL && K_PARALLEL == LBM_B200_PARALLEL
              && K_PERIODIC == LBM_B200_PERIODIC, "kind enums of kernels.cuh and lbm_b200.h must agree");

namespace {

thread_local std::string g_error;

int fail(int code, const char* fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_error = buf;
    return code;
}

#define CU(call)                                                                          \
    do {                                                                                  \
        cudaError_t e_ = (call);                                                          \
        if (e_ != cudaSuccess)                                                            \
            return fail(e_ == cudaErrorMemoryAllocation ? LBM_B200_ENOMEM : LBM_B200_ECUDA, \
                        "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

#define TRY(expr)                 \
    do {                          \
        int rc_ = (expr);         \
        if (rc_ != 0) return rc_; \
    } while (0)

// host copy of compute_feq (collision.hpp:34-51); this translation unit's host
// code is built with -ffp-contract=off, so the association below is what runs.
void feq_host(int Q, double rho, const double* u, double* out)
{
    double vel[27 * 3], w[27];
    lbm_b200_model(Q, vel, w);
    for (int q = 0; q < Q; ++q) {
        double c_dot_u = 0.0, u_dot_u = 0.0;
        for (int d = 0; d < 3; ++d) {
            c_dot_u += vel[q * 3 + d] * u[d];
            u_dot_u += u[d] * u[d];
        }
        out[q] = w[q] * rho * (1 + c_dot_u / CS2 + (c_dot_u) * (c_dot_u) / TWO_CS4 - u_dot_u / TWO_CS2);
    }
}

template <typename F>
auto dispatch_q(int Q, F&& f)
{
    switch (Q) {
    case 15: return f(std::integral_constant<int, 15>{});
    case 19: return f(std::integral_constant<int, 19>{});
    default: return f(std::integral_constant<int, 27>{});
    }
}

} // namespace

// Handler maps of ONE lattice (the reference keeps a handler pointer in every Cell of each of its two
// Lattice_fields, cell.h:15).  layer[b] belongs to the physical buffer f[b].  Unless a "literal" edit made
// the lattices differ, both layers alias the same arrays.
struct GeoLayer {
    uint8_t* kind = nullptr;
    uint16_t* bcid = nullptr;
    uint32_t* mask = nullptr;      // link mask of the steps whose SOURCE is this buffer
    int xhint_lo = 0, xhint_hi = 0;   // SweepParams::xhint_* of those steps
    uint32_t* bits = nullptr;
    int* inplace = nullptr;        // cells BGK-collided in place when this buffer is the DESTINATION
    int n_inplace = 0;
    double solid = 0.0;            // fraction of the interior cells that are not streamed (steps whose SOURCE is this buffer)
};

struct lbm_b200 {
    int Q = 0;
    int device = 0;
    Layout g{};
    // The split ("slow") axis of a slab is z (axis 2, planes = x-y planes) or y (axis 1, planes = x-z planes);
    // zl_global / z_first are the global length of THAT axis and the global index of the slab's local plane 1.
    int axis = 2;
    int zl_global = 0;
    int z_first = 1;
    double tau = 1.0;
    int exact = 0;

    double* f[2] = { nullptr, nullptr };
    int cur = 0;               // f[cur] is the collide field
    GeoLayer layer[2];
    bool split = false;        // the two lattices carry different handlers
    BcRec* d_bc = nullptr;
    int d_bc_cap = 0;
    std::vector<lbm_b200_bc> h_bc;   // the handler table
    bool geom_dirty = true;
    bool null_tagged = false;        // unsplit layers carry set_nonfluid_cells_nullcollide tags ...
    uint64_t null_tagged_at = 0;     // ... made at this step count (the tagged lattice alternates with the swaps)
    unsigned int* d_counters = nullptr;   // scratch words for the geometry kernels

    double* d_rho = nullptr;       // persistent read-out staging
    double* d_u = nullptr;
    bool first = true;         // boundary cells hold host-visible (stored) values
    bool materialized = true;  // boundary cells of f[cur] hold the reference's values
    int wrap_z = 0;
    bool ring_lo = false, ring_hi = false;   // periodic z must be closed by a slab ring on that side

    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;      // device->host copies of the split read-out
    // further copy streams (LBM_B200_COPY_STREAMS, default 2): the density and the velocity copies of a chunk travel on
    // different streams, i.e. on different copy engines.  A sweep that saturates the HBM starves a single engine's reads
    // (measured: 57 -> 28 GB/s once the sweep reached 99 % of the copy bandwidth), two engines get twice the share.
    static constexpr int MAX_COPY_STREAMS = 4;
    cudaStream_t copy_extra[MAX_COPY_STREAMS - 1] = {};
    cudaEvent_t ev_extra[MAX_COPY_STREAMS - 1] = {};
    int n_copy_streams = 2;
    cudaEvent_t ev_a = nullptr, ev_b = nullptr;
    static constexpr int READOUT_CHUNKS = 8;
    cudaEvent_t ev_chunk[READOUT_CHUNKS] = {};
    cudaEvent_t ev_copied = nullptr;
    bool readout_pending = false;
    bool timed = false;
    uint64_t steps = 0;
    uint64_t full_halo_at = ~0ull;   // time level for which neighbours pushed their full edge planes
    uint64_t launches = 0;
    uint64_t tma_launches = 0;       // how many of them were sweep_tma_kernel
    bool edges_done = false;   // split-phase state

    // CUDA graphs of step runs (graph_mode: 1 always, 0 never, -1 automatic)
    int graph_mode = -1;
    static constexpr int GRAPH_STEPS = 16;          // steps per graph (even: the buffer index returns)
    cudaGraphExec_t graph[2] = { nullptr, nullptr };   // [cur at entry]
    uint64_t graph_launches = 0;                    // kernel launches inside one graph
    uint64_t graph_tma_launches = 0;

    // direct peer stores: [side]
    double* peer_f[2][2] = { { nullptr, nullptr }, { nullptr, nullptr } };   // [side][buffer]
    long long peer_qstride[2] = { 0, 0 };
    long long peer_off[2] = { 0, 0 };
    void* peer_ipc_base[2] = { nullptr, nullptr };
    cudaIpcMemHandle_t peer_ipc_handle[2] = {};
    bool peer_ipc_shared = false;                   // both sides map the same exporter (ring of two)
    unsigned long long* d_flags = nullptr;          // [0],[1]: sweeps completed by the neighbour on that side,
                                                    // [2]: sweeps completed by this slab (the epoch), [4]: scratch
    unsigned long long* peer_flag[2] = { nullptr, nullptr };   // the neighbour's counter for us
    int* d_halo_error = nullptr;
    int* h_halo_error = nullptr;                    // pinned, mapped: set by the wait kernel on a time-out
    int* h_halo_error_dev = nullptr;                // its device address
    int clock_khz = 1965000;
    int sweep_mode = -1;                            // SWEEP_* for the interior launch; -1: chosen per geometry (LBM_B200_SWEEP_MODE)
    int xface_mode = 1;                             // 1: SWEEP_XFACE where the host has a guess for an x face (LBM_B200_XFACE=0: never)
    // TMA-fed sweep (sweep_tma_kernel): 1 where possible, 0 / -1 never (LBM_B200_TMA) -- an opt-in engine
    int tma_mode = -1;
    int tma_bx = 0;                                 // box width chosen for this lattice (0: no tensor maps)
    int sm_count = 148;
    CUtensorMap tmap[2][2];                         // [buffer][shifted]: 4-D views (x, y, z, q) of the two lattices,
                                                    // box bx x 256/bx, and (bx+2) x 256/bx for populations with c_x != 0
    long long pull_offset[27] = {};                 // c_z*plane + c_y*P + c_x per direction
    unsigned long long* d_trace = nullptr;         // LBM_B200_HALO_TRACE=<file prefix>: wait-kernel timestamps
    static constexpr int TRACE_EPOCHS = 8192;

    size_t ncell() const { return (size_t) (g.xl + 2) * (g.yl + 2) * (g.zl + 2); }
    size_t field_bytes() const { return (size_t) g.qstride * Q * sizeof(double); }
    size_t map_elems() const { return (size_t) g.qstride; }
    size_t bits_words() const { return map_elems() / 32 + 2; }
    bool lo_interface() const { return z_first != 1; }
    int nslow() const { return n_slow(g); }
    bool hi_interface() const { return z_first + nslow() - 1 != zl_global; }
    bool is_slab() const { return nslow() != zl_global; }
    int yl_global() const { return axis == 1 ? zl_global : g.yl; }      // logical global lengths
    int zl_global_z() const { return axis == 1 ? g.zl : zl_global; }
};

namespace {

// temporary device allocation that cannot leak on an early error return
struct DevBuf {
    void* p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    template <typename T> T* as() const { return static_cast<T*>(p); }
};

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev)
    {
        if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
        if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
    }
    ~DeviceGuard()
    {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

#define GUARD(h)                                                                     \
    if (!(h)) return fail(LBM_B200_EINVAL, "null handle");                           \
    DeviceGuard guard_((h)->device);                                                 \
    if (!guard_.ok) return fail(LBM_B200_ECUDA, "cannot select CUDA device %d", (h)->device)

dim3 map_grid(const Layout& g, int nz) { return dim3((g.xl + 2 + 127) / 128, g.yl + 2, nz); }

int fill_weights(lbm_b200* h, int buffer)
{
    dispatch_q(h->Q, [&](auto Qc) {
        fill_weights_kernel<decltype(Qc)::value><<<148 * 8, 256, 0, h->stream>>>(h->f[buffer], h->g.qstride);
        return 0;
    });
    h->launches++;
    CU(cudaGetLastError());
    return 0;
}

void drop_graphs(lbm_b200* h)
{
    for (int i = 0; i < 2; ++i) {
        if (h->graph[i]) cudaGraphExecDestroy(h->graph[i]);
        h->graph[i] = nullptr;
    }
}

void mark_geometry_dirty(lbm_b200* h)
{
    h->geom_dirty = true;
    h->materialized = false;
    drop_graphs(h);
}

// ---- the handler table ---------------------------------------------------------------------------
int sync_table(lbm_b200* h)
{
    const size_t n = std::max<size_t>(1, h->h_bc.size());
    std::vector<BcRec> recs(n);
    memset(recs.data(), 0, recs.size() * sizeof(BcRec));
    for (size_t i = 0; i < h->h_bc.size(); ++i) {
        recs[i].kind = h->h_bc[i].kind;
        for (int d = 0; d < 3; ++d) recs[i].v[d] = h->h_bc[i].v[d];
        recs[i].rho = h->h_bc[i].rho;
        if (recs[i].kind == LBM_B200_INFLOW) feq_host(h->Q, recs[i].rho, recs[i].v, recs[i].feq);
    }
    if ((int) n > h->d_bc_cap) {
        // kernels in flight may still read the old table
        CU(cudaStreamSynchronize(h->stream));
        if (h->d_bc) CU(cudaFree(h->d_bc));
        h->d_bc = nullptr;
        const size_t cap = std::max<size_t>(64, 2 * n);
        CU(cudaMalloc(&h->d_bc, cap * sizeof(BcRec)));
        h->d_bc_cap = (int) cap;
    }
    CU(cudaMemcpyAsync(h->d_bc, recs.data(), n * sizeof(BcRec), cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));   // recs goes out of scope
    drop_graphs(h);
    return 0;
}

bool known_kind(int k) { return k >= K_FLUID && k < K_COUNT; }

int check_table(const lbm_b200_bc* table, int n)
{
    if (n < 0 || n > 65535 || (n > 0 && !table)) return fail(LBM_B200_EINVAL, "bad handler table");
    for (int i = 0; i < n; ++i)
        if (!known_kind(table[i].kind)) return fail(LBM_B200_EINVAL, "handler %d: unknown kind %d", i, table[i].kind);
    return 0;
}

// ---- layers ----------------------------------------------------------------------------------------
int alloc_layer_maps(lbm_b200* h, GeoLayer& L)
{
    CU(cudaMalloc(&L.kind, h->map_elems()));
    CU(cudaMalloc(&L.bcid, h->map_elems() * sizeof(uint16_t)));
    CU(cudaMalloc(&L.mask, h->map_elems() * sizeof(uint32_t)));
    CU(cudaMalloc(&L.bits, h->bits_words() * sizeof(uint32_t)));
    return 0;
}

void free_layer(GeoLayer& L, bool maps)
{
    if (maps) {
        if (L.kind) cudaFree(L.kind);
        if (L.bcid) cudaFree(L.bcid);
        if (L.mask) cudaFree(L.mask);
        if (L.bits) cudaFree(L.bits);
    }
    if (L.inplace) cudaFree(L.inplace);
    L = GeoLayer();
}

// the two lattices start to carry different handlers
int ensure_split(lbm_b200* h)
{
    if (h->split) return 0;
    if (h->is_slab())
        return fail(LBM_B200_ESTATE, "handlers that differ between the two lattices (literal edits) are limited to whole domains, not slabs");
    CU(cudaStreamSynchronize(h->stream));
    if (h->layer[1].inplace && h->layer[1].inplace != h->layer[0].inplace) cudaFree(h->layer[1].inplace);
    h->layer[1] = GeoLayer();
    TRY(alloc_layer_maps(h, h->layer[1]));
    CU(cudaMemcpyAsync(h->layer[1].kind, h->layer[0].kind, h->map_elems(), cudaMemcpyDeviceToDevice, h->stream));
    CU(cudaMemcpyAsync(h->layer[1].bcid, h->layer[0].bcid, h->map_elems() * sizeof(uint16_t), cudaMemcpyDeviceToDevice, h->stream));
    h->split = true;
    if (h->null_tagged) {
        // the tags sit in the lattice that was the collide field when they were made
        const int tagged = ((h->steps - h->null_tagged_at) & 1) ? 1 - h->cur : h->cur;
        const long long n = (long long) h->map_elems();
        untag_null_kernel<<<(unsigned) ((n + 255) / 256), 256, 0, h->stream>>>(h->layer[1 - tagged].kind, h->layer[1 - tagged].bcid,
                                                                              h->d_bc, (int) h->h_bc.size(), n);
        h->launches++;
        CU(cudaGetLastError());
        h->null_tagged = false;
    }
    mark_geometry_dirty(h);
    return 0;
}

// both lattices carry the same handlers again (set_geometry replaces everything)
void unsplit(lbm_b200* h)
{
    if (!h->split) return;
    cudaStreamSynchronize(h->stream);
    free_layer(h->layer[1], true);
    h->layer[1] = h->layer[0];
    h->split = false;
}

// which layers an edit touches: both lattices, or (literal) the collide field's only
struct LayerSel {
    GeoLayer* l[2];
    int n;
};
int select_layers(lbm_b200* h, int literal, LayerSel* sel)
{
    if (literal) {
        TRY(ensure_split(h));
        sel->l[0] = &h->layer[h->cur];
        sel->n = 1;
    } else {
        sel->l[0] = &h->layer[0];
        sel->l[1] = &h->layer[1];
        sel->n = h->split ? 2 : 1;
    }
    return 0;
}

// builds link mask / bit map of the steps src -> dst and the list of cells collided in place in dst
int build_step_maps(lbm_b200* h, int src, int dst, bool* periodic_z)
{
    const Layout& g = h->g;
    GeoLayer& S = h->layer[src];
    GeoLayer& D = h->layer[dst];
    CU(cudaMemsetAsync(S.mask, 0x80, h->map_elems() * sizeof(uint32_t), h->stream));
    CU(cudaMemsetAsync(S.bits, 0, h->bits_words() * sizeof(uint32_t), h->stream));
    CU(cudaMemsetAsync(h->d_counters, 0, 4 * sizeof(unsigned int), h->stream));
    CU(cudaMemsetAsync(h->d_counters + 8, 0, 4 * sizeof(unsigned int), h->stream));
    const int lo = h->lo_interface(), hi = h->hi_interface();
    dispatch_q(h->Q, [&](auto Qc) {
        build_mask_kernel<decltype(Qc)::value><<<map_grid(g, g.zl + 2), 128, 0, h->stream>>>(S.kind, D.kind, S.bcid, S.mask, S.bits, g, lo, hi, h->d_counters);
        return 0;
    });
    h->launches++;
    CU(cudaGetLastError());
    unsigned int c[12];
    CU(cudaMemcpyAsync(c, h->d_counters, sizeof c, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    // what the cells of the two x faces mostly see next to them (SweepParams::xhint_*; only a guess that the
    // kernel checks against each cell's link mask)
    const long long face_cells = (long long) std::max(g.yl - 2, 0) * std::max(g.zl - 2, 0);
    for (int side = 0; side < 2; ++side) {
        const unsigned int noslip = c[8 + 2 * side], periodic = c[9 + 2 * side];
        int hint = 0;
        if (g.xl >= 2 && face_cells > 0) {
            if (2ll * noslip > face_cells) hint = 1;
            else if (2ll * periodic > face_cells) hint = 2;
        }
        (side == 0 ? S.xhint_lo : S.xhint_hi) = hint;
    }
    if (c[1] & 1u) *periodic_z = true;
    if (D.inplace) CU(cudaFree(D.inplace));
    D.inplace = nullptr;
    D.n_inplace = (int) c[0];
    if (D.n_inplace) {
        CU(cudaMalloc(&D.inplace, (size_t) D.n_inplace * sizeof(int)));
        collect_inplace_kernel<<<map_grid(g, g.zl + 2), 128, 0, h->stream>>>(S.kind, D.kind, g, lo, hi, D.inplace, c[0], h->d_counters + 2);
        h->launches++;
        CU(cudaGetLastError());
        CU(cudaStreamSynchronize(h->stream));
    }
    S.solid = (double) c[3] / ((double) g.xl * g.yl * g.zl);
    return 0;
}

// link masks, in-place lists, periodic-z bookkeeping
int commit_geometry(lbm_b200* h)
{
    if (!h->geom_dirty) return 0;
    const Layout& g = h->g;
    bool periodic_z = false;
    if (h->split) {
        TRY(build_step_maps(h, 0, 1, &periodic_z));
        TRY(build_step_maps(h, 1, 0, &periodic_z));
    } else {
        if (h->layer[1].inplace && h->layer[1].inplace != h->layer[0].inplace) cudaFree(h->layer[1].inplace);
        h->layer[1] = h->layer[0];
        h->layer[1].inplace = nullptr;
        TRY(build_step_maps(h, 0, 0, &periodic_z));
        h->layer[1] = h->layer[0];
    }
    h->ring_lo = h->ring_hi = false;
    if (periodic_z && h->is_slab()) {
        // closed by a ring of slabs instead: the first and the last slab must be each other's neighbours
        h->ring_lo = h->z_first == 1;
        h->ring_hi = h->z_first + h->nslow() - 1 == h->zl_global;
        periodic_z = false;
    }
    h->wrap_z = periodic_z ? 1 : 0;
    if (h->layer[0].n_inplace && h->is_slab())
        return fail(LBM_B200_EINVAL, "%d ghost-shell cells carry the fluid handler; on a multi-slab domain the "
                    "whole shell must be covered by boundary conditions", h->layer[0].n_inplace);
    h->geom_dirty = false;
    return 0;
}

void fill_sweep_params(lbm_b200* h, SweepParams& p, int z0, bool with_peers, int z_step, int* grid_xy)
{
    const Layout& g = h->g;
    const GeoLayer& S = h->layer[h->cur];
    p.src = h->f[h->cur];
    p.dst = h->f[1 - h->cur];
    p.mask = S.mask;
    p.bits = S.bits;
    p.kind = S.kind;
    p.bcid = S.bcid;
    p.bc = h->d_bc;
    p.g = g;
    p.z0 = z0;
    p.z_step = z_step;
    int tshift = 0;
    while ((1 << tshift) < LBM_SWEEP_THREADS) ++tshift;
    int shift = tshift;                  // all threads along x ...
    while (shift > 5 && (1 << (shift - 1)) >= g.xl) --shift;   // ... unless the row is short
    p.bx_shift = shift;
    p.first = h->first ? 1 : 0;
    p.xhint_lo = h->first ? 0 : S.xhint_lo;     // (first step: wall values are the stored ones, nothing to guess)
    p.xhint_hi = h->first ? 0 : S.xhint_hi;
    p.wrap_z = h->wrap_z;
    p.tau = h->tau;
    p.omega = 1.0 / h->tau;
    if (with_peers) {
        const int dstbuf = 1 - h->cur;
        p.up_dst = h->peer_f[LBM_B200_UP][dstbuf];
        p.up_qstride = h->peer_qstride[LBM_B200_UP];
        p.up_off = h->peer_off[LBM_B200_UP];
        p.dn_dst = h->peer_f[LBM_B200_DOWN][dstbuf];
        p.dn_qstride = h->peer_qstride[LBM_B200_DOWN];
        p.dn_off = h->peer_off[LBM_B200_DOWN];
    }
    for (int q = 0; q < h->Q; ++q) {
        p.srcq[q] = p.src + (long long) q * g.qstride - h->pull_offset[q];
        p.dstq[q] = p.dst + (long long) q * g.qstride;
    }
    const int bx = 1 << shift, by = LBM_SWEEP_THREADS >> shift;
    grid_xy[0] = (g.xl + bx - 1) / bx;
    grid_xy[1] = (n_mid(g) + by - 1) / by;
}

int launch_sweep(lbm_b200* h, int z0, int nz, bool with_peers, int z_step, int mode)
{
    if (nz <= 0) return 0;
    SweepParams p{};
    int gxy[2];
    fill_sweep_params(h, p, z0, with_peers, z_step, gxy);
    dim3 grid(gxy[0], gxy[1], nz);
    dispatch_q(h->Q, [&](auto Qc) {
        constexpr int Q = decltype(Qc)::value;
        auto go = [&](auto Ex) {
            constexpr bool EX = decltype(Ex)::value;
            // x-face guesses only where the host has one for this source layer (and never on the first step)
            const bool xface = mode == SWEEP_SPECULATIVE && h->xface_mode != 0 && (p.xhint_lo | p.xhint_hi) != 0;
            if (xface) sweep_kernel<Q, EX, SWEEP_XFACE><<<grid, LBM_SWEEP_THREADS, 0, h->stream>>>(p);
            else if (mode == SWEEP_SPLIT) sweep_kernel<Q, EX, SWEEP_SPLIT><<<grid, LBM_SWEEP_THREADS, 0, h->stream>>>(p);
            else if (mode == SWEEP_CHECKED) sweep_kernel<Q, EX, SWEEP_CHECKED><<<grid, LBM_SWEEP_THREADS, 0, h->stream>>>(p);
            else sweep_kernel<Q, EX, SWEEP_SPECULATIVE><<<grid, LBM_SWEEP_THREADS, 0, h->stream>>>(p);
        };
        if (h->exact) go(std::true_type{});
        else go(std::false_type{});
        return 0;
    });
    h->launches++;
    CU(cudaGetLastError());
    return 0;
}

// cuTensorMapEncodeTiled through the runtime (no link dependency on libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn tensor_map_encoder()
{
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
            cudaGetLastError();
            p = nullptr;
        }
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

// 4-D tensor maps (element x of a padded row, y, z, q) over both lattices; boxes of bx (and bx+2) x 256/bx x 1 x 1 doubles
int make_tensor_maps(lbm_b200* h)
{
    const Layout& g = h->g;
    h->tma_bx = 0;
    EncodeTiledFn enc = tensor_map_encoder();
    if (!enc) return 0;                         // stays on sweep_kernel
    const int bx = g.xl >= 112 ? 128 : 32;          // narrow lattices take narrower, taller boxes
    if (g.swap || g.P < bx + 2 || g.yl + 2 < 256 / bx) return 0;      // (z-slabs only)
    static const int promo = [] { const char* e = getenv("LBM_B200_TMA_L2PROMO"); return e ? atoi(e) : 0; }();
    for (int b = 0; b < 4; ++b) {
        const cuuint64_t dims[4] = { (cuuint64_t) g.P, (cuuint64_t) g.yl + 2, (cuuint64_t) g.zl + 2, (cuuint64_t) h->Q };
        const cuuint64_t strides[3] = { (cuuint64_t) g.P * 8, (cuuint64_t) g.plane * 8, (cuuint64_t) g.qstride * 8 };
        const cuuint32_t box[4] = { (cuuint32_t) (bx + 2 * (b & 1)), (cuuint32_t) (256 / bx), 1, 1 };
        const cuuint32_t estr[4] = { 1, 1, 1, 1 };
        const CUresult r = enc(&h->tmap[b >> 1][b & 1], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, h->f[b >> 1] + TMA_X0, dims, strides, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                               promo == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : (promo == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_NONE),
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return 0;
    }
    h->tma_bx = bx;
    return 0;
}

// does this launch go through the TMA-fed kernel?
bool tma_wanted(const lbm_b200* h, int nz, int mode)
{
    if (!h->tma_bx || h->tma_mode == 0 || h->first || mode != SWEEP_SPECULATIVE || nz <= 0) return false;
    // opt-in only: measured at 0.42-0.51 of the HBM roofline against 0.946 for sweep_kernel, because 16-24 consumer
    // warps cannot collide and store as many cells per second as 32 resident warps (profiles/variants_r07_tma.txt)
    return h->tma_mode > 0;
}

template <int Q, bool EX, int BX>
int launch_tma_one(lbm_b200* h, const SweepParams& p, int nz)
{
    using C = TmaCfg<Q>;
    const Layout& g = h->g;
    const int tiles_x = (g.xl + BX - 1) / BX, tiles_y = (g.yl + C::CELLS / BX - 1) / (C::CELLS / BX);
    const long long n_tiles = (long long) tiles_x * tiles_y * nz;
    static bool attr_set = false;     // per instantiation
    if (!attr_set) {
        CU(cudaFuncSetAttribute(sweep_tma_kernel<Q, EX, BX>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
        attr_set = true;
    }
    const int blocks = (int) std::min<long long>(n_tiles, h->sm_count);
    sweep_tma_kernel<Q, EX, BX><<<blocks, C::THREADS, C::SMEM_BYTES, h->stream>>>(p, h->tmap[h->cur][0], h->tmap[h->cur][1], tiles_x, tiles_y, nz);
    return 0;
}

int launch_sweep_tma(lbm_b200* h, int z0, int nz, bool with_peers)
{
    SweepParams p{};
    int gxy[2];
    fill_sweep_params(h, p, z0, with_peers, 1, gxy);
    int rc = dispatch_q(h->Q, [&](auto Qc) {
        constexpr int Q = decltype(Qc)::value;
        auto go = [&](auto Ex) {
            constexpr bool EX = decltype(Ex)::value;
            return h->tma_bx == 128 ? launch_tma_one<Q, EX, 128>(h, p, nz) : launch_tma_one<Q, EX, 32>(h, p, nz);
        };
        return h->exact ? go(std::true_type{}) : go(std::false_type{});
    });
    TRY(rc);
    h->launches++;
    h->tma_launches++;
    CU(cudaGetLastError());
    return 0;
}

// CUDA loads kernels lazily, on their first launch, and a load may have to wait for the device to drain.  Inside
// the time loop of a multi-slab run that is a deadlock: slab A's wait kernel spins for slab B's signal while the
// host is stuck loading the kernel variant slab B launches for the first time (seen with y-slabs whose solid
// fraction selects SWEEP_CHECKED for some slabs only; it ends with the hand-shake timeout).  So everything that
// can be launched between two synchronisations is loaded when the handle is created.
int preload_step_kernels(int Q)
{
    cudaFuncAttributes a;
    dispatch_q(Q, [&](auto Qc) {
        constexpr int QQ = decltype(Qc)::value;
        cudaFuncGetAttributes(&a, sweep_kernel<QQ, false, SWEEP_SPECULATIVE>);
        cudaFuncGetAttributes(&a, sweep_kernel<QQ, false, SWEEP_CHECKED>);
        cudaFuncGetAttributes(&a, sweep_kernel<QQ, false, SWEEP_SPLIT>);
        cudaFuncGetAttributes(&a, sweep_kernel<QQ, false, SWEEP_XFACE>);
        cudaFuncGetAttributes(&a, sweep_kernel<QQ, true, SWEEP_XFACE>);
        cudaFuncGetAttributes(&a, sweep_kernel<QQ, true, SWEEP_SPECULATIVE>);
        cudaFuncGetAttributes(&a, sweep_kernel<QQ, true, SWEEP_CHECKED>);
        cudaFuncGetAttributes(&a, sweep_kernel<QQ, true, SWEEP_SPLIT>);
        cudaFuncGetAttributes(&a, ghost_fluid_kernel<QQ, false>);
        cudaFuncGetAttributes(&a, ghost_fluid_kernel<QQ, true>);
        cudaFuncGetAttributes(&a, sweep_tma_kernel<QQ, false, 128>);
        cudaFuncGetAttributes(&a, sweep_tma_kernel<QQ, true, 128>);
        cudaFuncGetAttributes(&a, sweep_tma_kernel<QQ, false, 32>);
        cudaFuncGetAttributes(&a, sweep_tma_kernel<QQ, true, 32>);
        return 0;
    });
    cudaFuncGetAttributes(&a, halo_wait_kernel);
    cudaFuncGetAttributes(&a, halo_signal_kernel);
    CU(cudaGetLastError());
    return 0;
}

// how a launch over whole planes learns which cells are bulk cells, for the current source layer
int interior_mode(const lbm_b200* h)
{
    if (h->split) return SWEEP_SPLIT;          // MASK_NOCOLLIDE cells exist
    if (h->sweep_mode >= 0) return h->sweep_mode;
    // large solid regions: look at the bit before pulling (see SWEEP_CHECKED)
    return h->layer[h->cur].solid > 0.10 ? SWEEP_CHECKED : SWEEP_SPECULATIVE;
}

// planes [z0, z0 + nz)
int sweep_planes(lbm_b200* h, int z0, int nz, bool with_peers)
{
    const int mode = interior_mode(h);
    if (tma_wanted(h, nz, mode)) return launch_sweep_tma(h, z0, nz, with_peers);
    return launch_sweep(h, z0, nz, with_peers, 1, mode);
}

// cells that are collided without being streamed (K1g), in the buffer that becomes the collide field
int launch_inplace(lbm_b200* h)
{
    const GeoLayer& D = h->layer[1 - h->cur];
    if (!D.n_inplace) return 0;
    double* field = h->f[1 - h->cur];
    dispatch_q(h->Q, [&](auto Qc) {
        constexpr int Q = decltype(Qc)::value;
        const int blocks = (D.n_inplace + 127) / 128;
        if (h->exact) ghost_fluid_kernel<Q, true><<<blocks, 128, 0, h->stream>>>(field, h->g.qstride, D.inplace, D.n_inplace, h->tau, 1.0 / h->tau);
        else ghost_fluid_kernel<Q, false><<<blocks, 128, 0, h->stream>>>(field, h->g.qstride, D.inplace, D.n_inplace, h->tau, 1.0 / h->tau);
        return 0;
    });
    h->launches++;
    CU(cudaGetLastError());
    return 0;
}

bool has_peers(const lbm_b200* h) { return h->peer_flag[0] || h->peer_flag[1]; }

// see halo_wait_kernel / halo_signal_kernel
int halo_wait(lbm_b200* h)
{
    if (!has_peers(h)) return 0;
    // (the clock rate is read once at creation: cudaDeviceGetAttribute(cudaDevAttrClockRate) is a
    //  driver round trip of milliseconds and made the host the bottleneck when it ran every step)
    static const long long timeout_ms = [] { const char* e = getenv("LBM_B200_HALO_TIMEOUT_MS"); return e ? std::max(1LL, atoll(e)) : 20000LL; }();
    const long long timeout = (long long) h->clock_khz * timeout_ms;   // cycles; default ~20 s
    halo_wait_kernel<<<1, 1, 0, h->stream>>>(h->peer_flag[LBM_B200_DOWN] ? h->d_flags + LBM_B200_DOWN : nullptr,
                                             h->peer_flag[LBM_B200_UP] ? h->d_flags + LBM_B200_UP : nullptr,
                                             h->d_flags + 2, timeout, h->d_halo_error, h->h_halo_error_dev,
                                             h->d_trace, lbm_b200::TRACE_EPOCHS);
    h->launches++;
    CU(cudaGetLastError());
    return 0;
}
int halo_signal(lbm_b200* h)
{
    if (!has_peers(h)) return 0;
    halo_signal_kernel<<<1, 1, 0, h->stream>>>(h->peer_flag[LBM_B200_DOWN], h->peer_flag[LBM_B200_UP], h->d_flags + 2, h->d_flags + 4);
    h->launches++;
    CU(cudaGetLastError());
    return 0;
}
int halo_failed(const lbm_b200* h)
{
    if (h->h_halo_error && *(volatile int*) h->h_halo_error)
        return fail(LBM_B200_ETIMEOUT, "a neighbour slab did not complete its sweep within the hand-shake timeout; "
                    "the populations of this slab are no longer valid");
    return 0;
}
int halo_check(lbm_b200* h)
{
    if (!has_peers(h)) return halo_failed(h);
    int err = 0;
    CU(cudaMemcpyAsync(&err, h->d_halo_error, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    if (err) {
        if (h->h_halo_error) *h->h_halo_error = 1;
        return halo_failed(h);
    }
    return 0;
}

void finish_step(lbm_b200* h)
{
    h->cur = 1 - h->cur;      // Domain::swap, domain.hpp:169-172
    h->first = false;
    h->materialized = false;
    h->steps++;
}

// the launches of one time step on h->stream
int enqueue_step(lbm_b200* h)
{
    const int ns = h->nslow();
    if (has_peers(h) && ns >= 3) {
        // Only the two edge planes read ghost planes and feed the neighbours, so only they take
        // part in the hand-shake; the interior sweep that follows gives every neighbour a whole
        // step of slack before its next wait.
        TRY(halo_wait(h));
        TRY(launch_sweep(h, 1, 2, true, ns - 1, interior_mode(h)));
        TRY(halo_signal(h));
        TRY(sweep_planes(h, 2, ns - 2, false));
    } else {
        TRY(halo_wait(h));
        TRY(sweep_planes(h, 1, ns, true));
        TRY(halo_signal(h));
    }
    TRY(launch_inplace(h));
    finish_step(h);
    return 0;
}

bool graphs_wanted(const lbm_b200* h)
{
    if (h->graph_mode >= 0) return h->graph_mode != 0;
    // automatic: lattices whose sweep lasts a few microseconds, where the launch overhead shows
    return (size_t) h->g.xl * h->g.yl * h->g.zl <= ((size_t) 1 << 21);
}

// GRAPH_STEPS steps replayed from one CUDA graph (captured per buffer parity at entry)
int run_graph(lbm_b200* h)
{
    const int parity = h->cur;
    if (!h->graph[parity]) {
        const uint64_t launches0 = h->launches, steps0 = h->steps, tma0 = h->tma_launches;
        cudaGraph_t g = nullptr;
        CU(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
        int rc = 0;
        for (int s = 0; s < lbm_b200::GRAPH_STEPS && rc == 0; ++s) rc = enqueue_step(h);
        cudaError_t e = cudaStreamEndCapture(h->stream, &g);
        h->graph_launches = h->launches - launches0;
        h->graph_tma_launches = h->tma_launches - tma0;
        h->launches = launches0;      // nothing has run yet
        h->tma_launches = tma0;
        h->steps = steps0;
        if (rc != 0) { if (g) cudaGraphDestroy(g); return rc; }
        if (e != cudaSuccess) return fail(LBM_B200_ECUDA, "graph capture failed: %s", cudaGetErrorString(e));
        e = cudaGraphInstantiate(&h->graph[parity], g, 0);
        cudaGraphDestroy(g);
        if (e != cudaSuccess) return fail(LBM_B200_ECUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(e));
    }
    CU(cudaGraphLaunch(h->graph[parity], h->stream));
    h->launches += h->graph_launches;
    h->tma_launches += h->graph_tma_launches;
    h->steps += lbm_b200::GRAPH_STEPS;
    h->materialized = false;
    return 0;
}

// n steps on h->stream (no events, no state checks)
int enqueue_steps(lbm_b200* h, uint64_t n)
{
    while (n > 0) {
        if (!h->first && n >= (uint64_t) lbm_b200::GRAPH_STEPS && graphs_wanted(h)) {
            TRY(run_graph(h));
            n -= lbm_b200::GRAPH_STEPS;
        } else {
            TRY(enqueue_step(h));
            --n;
        }
    }
    return 0;
}

int ready_to_step(lbm_b200* h)
{
    if (h->edges_done) return fail(LBM_B200_ESTATE, "a split-phase step is in flight");
    TRY(halo_failed(h));
    TRY(commit_geometry(h));
    if ((h->ring_lo && !h->peer_f[LBM_B200_DOWN][0]) || (h->ring_hi && !h->peer_f[LBM_B200_UP][0]))
        return fail(LBM_B200_ESTATE, "periodic z on a multi-slab domain needs the first and last slab connected as a ring "
                    "(lbm_b200_connect / lbm_b200_connect_local on that side)");
    return 0;
}

int materialize(lbm_b200* h)
{
    TRY(commit_geometry(h));
    if (h->materialized) return 0;
    const Layout& g = h->g;
    const GeoLayer& L = h->layer[h->cur];
    // links from our boundary cells into an interface ghost plane can only be evaluated if the neighbours'
    // full planes were pushed there for this time level (lbm_b200_halo_push_all)
    const int lo_if = h->lo_interface(), hi_if = h->hi_interface();
    const int lo_open = lo_if && h->full_halo_at == h->steps;
    const int hi_open = hi_if && h->full_halo_at == h->steps;
    dispatch_q(h->Q, [&](auto Qc) {
        constexpr int Q = decltype(Qc)::value;
        if (h->exact) materialize_kernel<Q, true><<<map_grid(g, g.zl + 2), 128, 0, h->stream>>>(h->f[h->cur], L.kind, L.bcid, h->d_bc, g, lo_if, hi_if, lo_open, hi_open);
        else materialize_kernel<Q, false><<<map_grid(g, g.zl + 2), 128, 0, h->stream>>>(h->f[h->cur], L.kind, L.bcid, h->d_bc, g, lo_if, hi_if, lo_open, hi_open);
        return 0;
    });
    h->launches++;
    CU(cudaGetLastError());
    h->materialized = true;
    return 0;
}

// Geometry edits after time steps: first let the boundary cells of the OLD geometry take the values
// the reference holds (its non-fluid pass ran after every step), then remember that the next stream
// must pull stored values -- new boundary cells still carry their former fluid populations.
int before_geometry_change(lbm_b200* h)
{
    if (h->steps > 0 && !h->geom_dirty) TRY(materialize(h));
    h->first = true;
    return 0;
}

// one inclusive box in GLOBAL indices -> the local cells it covers in the given layers
int paint_box(lbm_b200* h, const LayerSel& sel, const uint64_t* e, int kind, uint16_t id)
{
    const Layout& g = h->g;
    const long long off = h->z_first - 1;    // along the split axis: local index = global index - off
    const int a = h->axis == 1 ? 2 : 4;      // position of that axis' extent inside the box
    long long lo[3] = { (long long) e[0], (long long) e[2], (long long) e[4] }, hi[3] = { (long long) e[1], (long long) e[3], (long long) e[5] };
    lo[a / 2] = std::max<long long>((long long) e[a] - off, 0);
    hi[a / 2] = std::min<long long>((long long) e[a + 1] - off, h->nslow() + 1);
    if (lo[a / 2] > hi[a / 2]) return 0;
    const int nx = (int) (hi[0] - lo[0] + 1), ny = (int) (hi[1] - lo[1] + 1), nz = (int) (hi[2] - lo[2] + 1);
    const int threads = nx >= 128 ? 128 : 32;
    dim3 grid((nx + threads - 1) / threads, ny, nz);
    for (int l = 0; l < sel.n; ++l) {
        paint_box_kernel<<<grid, threads, 0, h->stream>>>(sel.l[l]->kind, sel.l[l]->bcid, g, (int) lo[0], (int) lo[1], (int) lo[2], nx, ny, (uint8_t) kind, id);
        h->launches++;
    }
    CU(cudaGetLastError());
    return 0;
}

int check_box(const lbm_b200* h, const uint64_t* e, int b)
{
    const Layout& g = h->g;
    // the reference asserts these (domain.hpp:180-181)
    if (!(e[1] >= e[0] && e[3] >= e[2] && e[5] >= e[4])) return fail(LBM_B200_EINVAL, "box %d: end before begin", b);
    if (!(e[1] < (uint64_t) g.xl + 2 && e[3] < (uint64_t) h->yl_global() + 2 && e[5] < (uint64_t) h->zl_global_z() + 2))
        return fail(LBM_B200_EINVAL, "box %d: extent outside the domain", b);
    return 0;
}

// dense kind / handler-id planes [z_begin, z_begin + z_count) from the host -> checked on the device -> maps
int upload_map_planes(lbm_b200* h, const uint8_t* kind, const uint16_t* bc_id, int z_begin, int z_count, const LayerSel& sel)
{
    const Layout& g = h->g;
    const size_t n = (size_t) (g.xl + 2) * (g.yl + 2) * z_count;
    DevBuf b8, b16, bres;
    CU(cudaMalloc(&b8.p, n));
    CU(cudaMalloc(&b16.p, n * sizeof(uint16_t)));
    CU(cudaMalloc(&bres.p, sizeof(unsigned long long)));
    CU(cudaMemcpyAsync(b8.p, kind, n, cudaMemcpyHostToDevice, h->stream));
    if (bc_id) CU(cudaMemcpyAsync(b16.p, bc_id, n * sizeof(uint16_t), cudaMemcpyHostToDevice, h->stream));
    else CU(cudaMemsetAsync(b16.p, 0, n * sizeof(uint16_t), h->stream));
    CU(cudaMemsetAsync(bres.p, 0xFF, sizeof(unsigned long long), h->stream));
    validate_dense_kernel<<<map_grid(g, z_count), 128, 0, h->stream>>>(b8.as<uint8_t>(), b16.as<uint16_t>(), h->d_bc, (int) h->h_bc.size(),
                                                                        g, z_begin, bres.as<unsigned long long>());
    h->launches++;
    CU(cudaGetLastError());
    unsigned long long res = 0;
    CU(cudaMemcpyAsync(&res, bres.p, sizeof res, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));     // the caller may release its arrays after this call
    if (res != ~0ull) {                       // nothing has been changed
        const unsigned long long at = res / 8;
        const int why = (int) (res % 8);
        const long long x = at % (g.xl + 2), y = (at / (g.xl + 2)) % (g.yl + 2), z = at / ((unsigned long long) (g.xl + 2) * (g.yl + 2));
        const char* msg[] = { "", "unknown kind", "bc id outside the table", "kind differs from table[bc id].kind", "PERIODIC is a ghost-shell kind" };
        return fail(LBM_B200_EINVAL, "cell (%lld,%lld,%lld): %s", x, y, z, msg[why]);
    }
    for (int l = 0; l < sel.n; ++l) {
        scatter_map_kernel<uint8_t><<<map_grid(g, z_count), 128, 0, h->stream>>>(b8.as<uint8_t>(), sel.l[l]->kind, g, z_begin);
        scatter_map_kernel<uint16_t><<<map_grid(g, z_count), 128, 0, h->stream>>>(b16.as<uint16_t>(), sel.l[l]->bcid, g, z_begin);
        h->launches += 2;
    }
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(h->stream));     // the staging buffers are released on return
    return 0;
}

// (xl, yl_all, zl_all): the GLOBAL domain; the handle owns planes [first, first + local) of `axis` (1 = y, 2 = z)
int create_common(lbm_b200_t** out, int Q, uint64_t xl, uint64_t yl_all, uint64_t zl_all, int axis, uint64_t z_first,
                  uint64_t zl_local, double tau, int device)
{
    if (!out) return fail(LBM_B200_EINVAL, "null output pointer");
    *out = nullptr;
    if (Q != 15 && Q != 19 && Q != 27) return fail(LBM_B200_EINVAL, "Q must be 15, 19 or 27 (got %d)", Q);
    if (axis != 1 && axis != 2) return fail(LBM_B200_EINVAL, "split axis must be 1 (y) or 2 (z)");
    const uint64_t zl_global = axis == 1 ? yl_all : zl_all;          // length of the split axis
    if (xl == 0 || yl_all == 0 || zl_all == 0 || zl_local == 0)
        return fail(LBM_B200_EINVAL, "domain lengths must be positive");
    if (z_first < 1 || z_first + zl_local - 1 > zl_global)
        return fail(LBM_B200_EINVAL, "slab [%llu, %llu] outside 1..%llu", (unsigned long long) z_first,
                    (unsigned long long) (z_first + zl_local - 1), (unsigned long long) zl_global);
    if (!(tau > 0.0)) return fail(LBM_B200_EINVAL, "tau must be positive (got %g)", tau);
    const uint64_t yl = axis == 1 ? zl_local : yl_all, zl = axis == 1 ? zl_all : zl_local;   // local lengths
    if (yl + 2 > 65535 || zl + 2 > 65535) return fail(LBM_B200_EINVAL, "yl and zl are limited to 65533");
    const uint64_t P = (xl + 2 + 15) / 16 * 16;
    const uint64_t plane = P * ((axis == 1 ? zl : yl) + 2);         // one plane of the split axis
    const uint64_t qstride = plane * (zl_local + 2) + 16;
    if (qstride >= (1ull << 31)) return fail(LBM_B200_EINVAL, "slab too large: %llu padded cells per population (limit 2^31)", (unsigned long long) qstride);

    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(LBM_B200_ECUDA, "no CUDA device available (there is no CPU fallback)");
    }
    if (device < 0) {
        if (cudaGetDevice(&device) != cudaSuccess) return fail(LBM_B200_ECUDA, "cudaGetDevice failed");
    }
    if (device >= ndev) return fail(LBM_B200_EINVAL, "device %d out of range (%d visible)", device, ndev);

    lbm_b200* h = new lbm_b200();
    h->Q = Q;
    h->device = device;
    h->axis = axis;
    h->g.xl = (int) xl; h->g.yl = (int) yl; h->g.zl = (int) zl;
    h->g.P = (int) P; h->g.plane = (int) plane; h->g.qstride = (long long) qstride;
    h->g.swap = axis == 1 ? 1 : 0;
    h->g.sy = axis == 1 ? (int) plane : (int) P;
    h->g.sz = axis == 1 ? (int) P : (int) plane;
    h->zl_global = (int) zl_global;
    h->z_first = (int) z_first;
    h->tau = tau;
    if (const char* e = getenv("LBM_B200_GRAPHS")) h->graph_mode = atoi(e);
    if (const char* e = getenv("LBM_B200_SWEEP_MODE")) h->sweep_mode = std::max(-1, std::min(1, atoi(e)));
    if (const char* e = getenv("LBM_B200_XFACE")) h->xface_mode = atoi(e) != 0;
    if (const char* e = getenv("LBM_B200_TMA")) h->tma_mode = std::max(-1, std::min(1, atoi(e)));
    {
        double vel[27 * 3];
        lbm_b200_model(Q, vel, nullptr);
        for (int q = 0; q < Q; ++q)
            h->pull_offset[q] = (long long) vel[3 * q + 2] * h->g.sz + (long long) vel[3 * q + 1] * h->g.sy + (long long) vel[3 * q];
    }
    *out = h;   // so that the caller can destroy on failure below
    DeviceGuard guard(device);
    if (!guard.ok) { lbm_b200_destroy(h); *out = nullptr; return fail(LBM_B200_ECUDA, "cannot select CUDA device %d", device); }

    auto bail = [&](int rc) { std::string keep = g_error; lbm_b200_destroy(h); *out = nullptr; g_error = keep; return rc; };
#define CUB(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
        fail(e_ == cudaErrorMemoryAllocation ? LBM_B200_ENOMEM : LBM_B200_ECUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); \
        cudaGetLastError(); return bail(e_ == cudaErrorMemoryAllocation ? LBM_B200_ENOMEM : LBM_B200_ECUDA); } } while (0)
    CUB(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
    h->stream = h->own_stream;
    CUB(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    if (const char* e = getenv("LBM_B200_COPY_STREAMS")) h->n_copy_streams = std::max(1, std::min((int) lbm_b200::MAX_COPY_STREAMS, atoi(e)));
    for (int k = 0; k + 1 < h->n_copy_streams; ++k) {
        CUB(cudaStreamCreateWithFlags(&h->copy_extra[k], cudaStreamNonBlocking));
        CUB(cudaEventCreateWithFlags(&h->ev_extra[k], cudaEventDisableTiming));
    }
    CUB(cudaEventCreate(&h->ev_a));
    CUB(cudaEventCreate(&h->ev_b));
    for (auto& e : h->ev_chunk) CUB(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    CUB(cudaEventCreateWithFlags(&h->ev_copied, cudaEventDisableTiming));
    // one allocation for both lattices: a neighbour maps it with a single IPC handle
    // (IPC handles address whole allocations, so the hand-shake counters live in its tail)
    CUB(cudaMalloc(&h->f[0], 2 * h->field_bytes() + 256));
    h->f[1] = h->f[0] + (size_t) h->g.qstride * Q;
    h->d_flags = reinterpret_cast<unsigned long long*>(h->f[0] + 2 * (size_t) h->g.qstride * Q);
    CUB(cudaMemset(h->d_flags, 0, 256));
    if (alloc_layer_maps(h, h->layer[0]) != 0) return bail(LBM_B200_ENOMEM);
    h->layer[1] = h->layer[0];
    CUB(cudaMalloc(&h->d_counters, 16 * sizeof(unsigned int)));
    if (cudaDeviceGetAttribute(&h->clock_khz, cudaDevAttrClockRate, device) != cudaSuccess || h->clock_khz <= 0) {
        cudaGetLastError();
        h->clock_khz = 1965000;
    }
    if (cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || h->sm_count <= 0) {
        cudaGetLastError();
        h->sm_count = 148;
    }
    make_tensor_maps(h);
    if (preload_step_kernels(Q) != 0) return bail(LBM_B200_ECUDA);
    CUB(cudaMalloc(&h->d_halo_error, sizeof(int)));
    CUB(cudaMemset(h->d_halo_error, 0, sizeof(int)));
    CUB(cudaHostAlloc(&h->h_halo_error, sizeof(int), cudaHostAllocMapped | cudaHostAllocPortable));
    *h->h_halo_error = 0;
    CUB(cudaHostGetDevicePointer(&h->h_halo_error_dev, h->h_halo_error, 0));
    if (getenv("LBM_B200_HALO_TRACE")) {
        CUB(cudaMalloc(&h->d_trace, 2 * lbm_b200::TRACE_EPOCHS * sizeof(unsigned long long)));
        CUB(cudaMemset(h->d_trace, 0, 2 * lbm_b200::TRACE_EPOCHS * sizeof(unsigned long long)));
    }
    // every cell -- ghost shell included -- starts with the fluid handler (domain.hpp:87-93); the padding
    // between rows is NULL so that nothing ever treats it as a cell
    CUB(cudaMemsetAsync(h->layer[0].kind, K_NULL, h->map_elems(), h->stream));
    CUB(cudaMemsetAsync(h->layer[0].bcid, 0, h->map_elems() * sizeof(uint16_t), h->stream));
#undef CUB
    {
        const uint64_t whole[6] = { 0, xl + 1, 0, yl_all + 1, 0, zl_all + 1 };
        LayerSel sel{ { &h->layer[0], nullptr }, 1 };
        if (paint_box(h, sel, whole, K_FLUID, 0) != 0) return bail(LBM_B200_ECUDA);
    }
    lbm_b200_bc fluid{};
    fluid.kind = LBM_B200_FLUID;
    fluid.rho = 1.0;
    h->h_bc.assign(1, fluid);      // id 0: the fluid operator (cells of a fresh domain refer to it)
    if (sync_table(h) != 0) return bail(LBM_B200_ECUDA);
    if (fill_weights(h, 0) != 0 || fill_weights(h, 1) != 0) return bail(LBM_B200_ECUDA);
    if (cudaStreamSynchronize(h->stream) != cudaSuccess) { fail(LBM_B200_ECUDA, "initial fill failed"); return bail(LBM_B200_ECUDA); }
    return 0;
}

} // namespace

// ---------------------------------------------------------------------------
extern "C" {

const char* lbm_b200_last_error(void) { return g_error.c_str(); }
int lbm_b200_abi_version(void) { return LBM_B200_ABI_VERSION; }

int lbm_b200_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int lbm_b200_model(int Q, double* velocities, double* weights)
{
    if (Q != 15 && Q != 19 && Q != 27) return fail(LBM_B200_EINVAL, "Q must be 15, 19 or 27 (got %d)", Q);
    dispatch_q(Q, [&](auto Qc) {
        using L = Lattice<decltype(Qc)::value>;
        for (int q = 0; q < Q; ++q) {
            if (velocities) { velocities[3 * q] = L::cx(q); velocities[3 * q + 1] = L::cy(q); velocities[3 * q + 2] = L::cz(q); }
            if (weights) weights[q] = L::w(q);
        }
        return 0;
    });
    return 0;
}
int lbm_b200_model_inv(int Q, int q) { return Q - 1 - q; }
int lbm_b200_model_velocity_index(int Q, int u, int v, int w)
{
    if (Q != 15 && Q != 19 && Q != 27) return fail(LBM_B200_EINVAL, "Q must be 15, 19 or 27 (got %d)", Q);
    if (u < -1 || u > 1 || v < -1 || v > 1 || w < -1 || w > 1) return fail(LBM_B200_EINVAL, "components must be -1, 0 or 1");
    return dispatch_q(Q, [&](auto Qc) { return Lattice<decltype(Qc)::value>::index_of(u, v, w); });
}

int lbm_b200_create(lbm_b200_t** h, int Q, uint64_t xl, uint64_t yl, uint64_t zl, double tau, int device)
{
    return create_common(h, Q, xl, yl, zl, 2, 1, zl, tau, device);
}
int lbm_b200_create_slab(lbm_b200_t** h, int Q, uint64_t xl, uint64_t yl, uint64_t zl_global, uint64_t z_first,
                         uint64_t zl_local, double tau, int device)
{
    return create_common(h, Q, xl, yl, zl_global, 2, z_first, zl_local, tau, device);
}
int lbm_b200_create_slab_axis(lbm_b200_t** h, int Q, uint64_t xl, uint64_t yl, uint64_t zl, int axis, uint64_t first,
                              uint64_t local, double tau, int device)
{
    return create_common(h, Q, xl, yl, zl, axis, first, local, tau, device);
}

int lbm_b200_destroy(lbm_b200_t* h)
{
    if (!h) return 0;
    DeviceGuard guard(h->device);
    if (h->readout_pending && h->ev_copied) cudaEventSynchronize(h->ev_copied);
    if (h->own_stream) cudaStreamSynchronize(h->own_stream);
    if (h->d_trace) {   // diagnostic dump: "<prefix>.<device>.z<first plane>" with entry/exit ns per epoch
        std::vector<unsigned long long> t(2 * lbm_b200::TRACE_EPOCHS);
        cudaDeviceSynchronize();
        unsigned long long epochs = 0;
        cudaMemcpy(&epochs, h->d_flags + 2, sizeof epochs, cudaMemcpyDeviceToHost);
        if (cudaMemcpy(t.data(), h->d_trace, t.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost) == cudaSuccess) {
            const std::string path = std::string(getenv("LBM_B200_HALO_TRACE") ? getenv("LBM_B200_HALO_TRACE") : "halo_trace")
                    + "." + std::to_string(h->device) + ".z" + std::to_string(h->z_first);
            if (FILE* fp = fopen(path.c_str(), "w")) {
                for (unsigned long long e = 0; e < epochs && e < (unsigned long long) lbm_b200::TRACE_EPOCHS; ++e)
                    fprintf(fp, "%llu %llu %llu\n", e, t[2 * e], t[2 * e + 1]);
                fclose(fp);
            }
        }
        cudaFree(h->d_trace);
    }
    drop_graphs(h);
    if (h->peer_ipc_shared) h->peer_ipc_base[1] = nullptr;
    for (int s = 0; s < 2; ++s) {
        if (h->peer_ipc_base[s]) cudaIpcCloseMemHandle(h->peer_ipc_base[s]);
    }
    if (h->d_halo_error) cudaFree(h->d_halo_error);
    if (h->h_halo_error) cudaFreeHost(h->h_halo_error);
    if (h->f[0]) cudaFree(h->f[0]);
    if (h->layer[1].inplace && h->layer[1].inplace != h->layer[0].inplace) cudaFree(h->layer[1].inplace);
    h->layer[1].inplace = nullptr;
    if (h->split) free_layer(h->layer[1], true);
    free_layer(h->layer[0], true);
    if (h->d_bc) cudaFree(h->d_bc);
    if (h->d_counters) cudaFree(h->d_counters);
    if (h->d_rho) cudaFree(h->d_rho);
    if (h->d_u) cudaFree(h->d_u);
    if (h->ev_a) cudaEventDestroy(h->ev_a);
    if (h->ev_b) cudaEventDestroy(h->ev_b);
    for (auto& e : h->ev_chunk) if (e) cudaEventDestroy(e);
    if (h->ev_copied) cudaEventDestroy(h->ev_copied);
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    for (auto& st : h->copy_extra) if (st) cudaStreamDestroy(st);
    for (auto& e : h->ev_extra) if (e) cudaEventDestroy(e);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    cudaGetLastError();
    delete h;
    return 0;
}

int lbm_b200_set_arithmetic(lbm_b200_t* h, int mode)
{
    if (!h) return fail(LBM_B200_EINVAL, "null handle");
    if (mode != LBM_B200_FAST && mode != LBM_B200_EXACT) return fail(LBM_B200_EINVAL, "unknown arithmetic mode %d", mode);
    h->exact = mode == LBM_B200_EXACT;
    drop_graphs(h);
    return 0;
}
int lbm_b200_set_tau(lbm_b200_t* h, double tau)
{
    if (!h) return fail(LBM_B200_EINVAL, "null handle");
    if (!(tau > 0.0)) return fail(LBM_B200_EINVAL, "tau must be positive (got %g)", tau);
    h->tau = tau;
    drop_graphs(h);
    return 0;
}
int lbm_b200_set_graphs(lbm_b200_t* h, int mode)
{
    if (!h) return fail(LBM_B200_EINVAL, "null handle");
    h->graph_mode = mode < 0 ? -1 : (mode ? 1 : 0);
    return 0;
}
int lbm_b200_set_sweep_engine(lbm_b200_t* h, int tma, int checked)
{
    if (!h) return fail(LBM_B200_EINVAL, "null handle");
    h->tma_mode = tma < 0 ? -1 : (tma ? 1 : 0);
    h->sweep_mode = checked < 0 ? -1 : (checked ? SWEEP_CHECKED : SWEEP_SPECULATIVE);
    drop_graphs(h);
    return 0;
}
int lbm_b200_set_stream(lbm_b200_t* h, void* cuda_stream)
{
    GUARD(h);
    CU(cudaStreamSynchronize(h->stream));
    h->stream = cuda_stream ? (cudaStream_t) cuda_stream : h->own_stream;
    return 0;
}

int lbm_b200_set_handlers(lbm_b200_t* h, const lbm_b200_bc* table, int n_table)
{
    GUARD(h);
    TRY(check_table(table, n_table));
    for (size_t i = 0; i < h->h_bc.size() && i < (size_t) n_table; ++i)
        if (h->h_bc[i].kind != table[i].kind)
            return fail(LBM_B200_EINVAL, "handler %zu changes its kind (%d -> %d): a new table must extend the old one", i, h->h_bc[i].kind, table[i].kind);
    if ((size_t) n_table < h->h_bc.size()) return fail(LBM_B200_EINVAL, "a new handler table must extend the old one (%zu handlers)", h->h_bc.size());
    h->h_bc.assign(table, table + n_table);
    return sync_table(h);
}

int lbm_b200_paint_boxes(lbm_b200_t* h, const uint64_t* boxes6, const uint16_t* ids, int n)
{
    GUARD(h);
    if (n < 0 || (n > 0 && (!boxes6 || !ids))) return fail(LBM_B200_EINVAL, "bad box list");
    for (int b = 0; b < n; ++b) {
        TRY(check_box(h, boxes6 + 6 * b, b));
        if (ids[b] >= h->h_bc.size()) return fail(LBM_B200_EINVAL, "box %d: handler id %d outside the table (%zu handlers)", b, ids[b], h->h_bc.size());
    }
    if (n == 0) return 0;
    TRY(before_geometry_change(h));
    LayerSel sel;
    TRY(select_layers(h, 0, &sel));
    for (int b = 0; b < n; ++b) TRY(paint_box(h, sel, boxes6 + 6 * b, h->h_bc[ids[b]].kind, ids[b]));
    mark_geometry_dirty(h);
    return 0;
}

int lbm_b200_set_boxes(lbm_b200_t* h, const uint64_t* boxes6, const lbm_b200_bc* table, int n)
{
    GUARD(h);
    if (n < 0 || (n > 0 && (!boxes6 || !table))) return fail(LBM_B200_EINVAL, "bad box list");
    TRY(check_table(table, n));
    for (int b = 0; b < n; ++b) TRY(check_box(h, boxes6 + 6 * b, b));
    if (h->h_bc.size() + (size_t) n > 65535) return fail(LBM_B200_EINVAL, "more than 65535 boundary handlers");
    if (n == 0) return 0;
    std::vector<uint16_t> ids(n);
    for (int b = 0; b < n; ++b) {
        ids[b] = (uint16_t) h->h_bc.size();
        h->h_bc.push_back(table[b]);
    }
    TRY(sync_table(h));
    return lbm_b200_paint_boxes(h, boxes6, ids.data(), n);
}

int lbm_b200_set_geometry(lbm_b200_t* h, const uint8_t* kind, const uint16_t* bc_id, const lbm_b200_bc* table, int n_table)
{
    GUARD(h);
    if (!kind) return fail(LBM_B200_EINVAL, "kind map is null");
    TRY(check_table(table, n_table));
    if (!bc_id && n_table > 1) return fail(LBM_B200_EINVAL, "bc_id map required for a table of %d handlers", n_table);
    TRY(before_geometry_change(h));
    // the maps are checked against the NEW table; on failure the old table comes back
    std::vector<lbm_b200_bc> old_table = h->h_bc;
    h->h_bc.assign(table, table + n_table);
    TRY(sync_table(h));
    unsplit(h);
    LayerSel sel{ { &h->layer[0], nullptr }, 1 };
    const int rc = upload_map_planes(h, kind, bc_id, 0, h->g.zl + 2, sel);
    if (rc != 0) {
        const std::string keep = g_error;
        h->h_bc = old_table;
        sync_table(h);
        g_error = keep;
        return rc;
    }
    h->null_tagged = false;
    mark_geometry_dirty(h);
    return 0;
}

int lbm_b200_set_geometry_planes(lbm_b200_t* h, const uint8_t* kind, const uint16_t* bc_id, uint64_t z_begin, uint64_t z_count, int literal)
{
    GUARD(h);
    if (!kind || !bc_id) return fail(LBM_B200_EINVAL, "kind / bc_id map is null");
    if (z_begin + z_count > (uint64_t) h->g.zl + 2) return fail(LBM_B200_EINVAL, "plane range outside the slab");
    if (z_count == 0) return 0;
    TRY(before_geometry_change(h));
    LayerSel sel;
    TRY(select_layers(h, literal, &sel));
    TRY(upload_map_planes(h, kind, bc_id, (int) z_begin, (int) z_count, sel));
    mark_geometry_dirty(h);
    return 0;
}

int lbm_b200_get_geometry_planes(lbm_b200_t* h, uint8_t* kind, uint16_t* bc_id, uint64_t z_begin, uint64_t z_count)
{
    GUARD(h);
    if (z_begin + z_count > (uint64_t) h->g.zl + 2) return fail(LBM_B200_EINVAL, "plane range outside the slab");
    if (z_count == 0 || (!kind && !bc_id)) return 0;
    const Layout& g = h->g;
    const size_t n = (size_t) (g.xl + 2) * (g.yl + 2) * z_count;
    DevBuf b8, b16;
    if (kind) CU(cudaMalloc(&b8.p, n));
    if (bc_id) CU(cudaMalloc(&b16.p, n * sizeof(uint16_t)));
    const GeoLayer& L = h->layer[h->cur];
    // unsplit layers: the tags of set_nonfluid_cells_nullcollide sit in one lattice only (see tag_null_cells)
    const int former = !h->split && h->null_tagged && ((h->steps - h->null_tagged_at) & 1);
    gather_maps_kernel<<<map_grid(g, (int) z_count), 128, 0, h->stream>>>(L.kind, L.bcid, h->d_bc, (int) h->h_bc.size(), b8.as<uint8_t>(),
                                                                         b16.as<uint16_t>(), g, (int) z_begin, former);
    h->launches++;
    CU(cudaGetLastError());
    if (kind) CU(cudaMemcpyAsync(kind, b8.p, n, cudaMemcpyDeviceToHost, h->stream));
    if (bc_id) CU(cudaMemcpyAsync(bc_id, b16.p, n * sizeof(uint16_t), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return 0;
}

int lbm_b200_get_kind(lbm_b200_t* h, uint8_t* kind)
{
    if (!h || !kind) return fail(LBM_B200_EINVAL, "null argument");
    return lbm_b200_get_geometry_planes(h, kind, nullptr, 0, (uint64_t) h->g.zl + 2);
}

// mask: x-y planes of interior cells for the GLOBAL planes [mask_z_first, mask_z_first + mask_nz); cells whose
// mask value is 0 take handler `id` of the table (id < 0: a NoSlipBoundary is appended for them)
static int apply_fluid_mask(lbm_b200* h, const uint8_t* mask, int mask_z_first, int mask_nz, int literal, int id)
{
    if (!mask) return fail(LBM_B200_EINVAL, "mask is null");
    if (id >= (int) h->h_bc.size()) return fail(LBM_B200_EINVAL, "handler id %d outside the table (%zu handlers)", id, h->h_bc.size());
    if (id < 0 && h->h_bc.size() >= 65535) return fail(LBM_B200_EINVAL, "more than 65535 boundary handlers");
    const Layout& g = h->g;
    // local planes that are interior planes of the global domain and covered by the mask: own planes plus
    // the replicas of the neighbours' edge planes
    const int zoff = h->z_first - 1;                       // local z = global z - zoff
    const int lo = std::max(std::max(mask_z_first - zoff, 1 - zoff), 0);
    const int hi = std::min(std::min(mask_z_first + mask_nz - 1 - zoff, h->zl_global - zoff), h->nslow() + 1);
    if (lo > hi) return 0;
    TRY(before_geometry_change(h));
    LayerSel sel;
    TRY(select_layers(h, literal, &sel));
    if (id < 0) {
        lbm_b200_bc solid{};
        solid.kind = LBM_B200_NOSLIP;      // read_vtk_point_file<M, NoSlipBoundary<M>> (io/scenario.h:161-162)
        solid.rho = 1.0;
        id = (int) h->h_bc.size();
        h->h_bc.push_back(solid);
        TRY(sync_table(h));
    }
    DevBuf buf;
    dim3 grid;
    int y_lo, z_lo, mask_rows, y_shift, z_shift;
    if (h->axis == 2) {
        // mask planes are x-y planes: ship the planes [lo, hi] only
        const size_t plane_cells = (size_t) g.xl * g.yl;
        const int nz = hi - lo + 1;
        const int first_mask_plane = lo + zoff - mask_z_first;     // index into `mask`
        CU(cudaMalloc(&buf.p, plane_cells * nz));
        CU(cudaMemcpyAsync(buf.p, mask + plane_cells * first_mask_plane, plane_cells * nz, cudaMemcpyHostToDevice, h->stream));
        grid = dim3((g.xl + 127) / 128, g.yl, nz);
        y_lo = 1; z_lo = lo; mask_rows = g.yl; y_shift = 1; z_shift = lo;
    } else {
        // y-slab: the rows [lo, hi] of every x-y plane of the mask (mask_nz rows per plane); the mask is bytes, ship it whole
        const size_t bytes = (size_t) g.xl * mask_nz * g.zl;
        CU(cudaMalloc(&buf.p, bytes));
        CU(cudaMemcpyAsync(buf.p, mask, bytes, cudaMemcpyHostToDevice, h->stream));
        grid = dim3((g.xl + 127) / 128, hi - lo + 1, g.zl);
        y_lo = lo; z_lo = 1; mask_rows = mask_nz; y_shift = mask_z_first - zoff; z_shift = 1;
    }
    for (int l = 0; l < sel.n; ++l) {
        paint_mask_kernel<<<grid, 128, 0, h->stream>>>(buf.as<uint8_t>(), sel.l[l]->kind, sel.l[l]->bcid, g, y_lo, z_lo, mask_rows,
                                                       y_shift, z_shift, (uint8_t) h->h_bc[id].kind, (uint16_t) id);
        h->launches++;
    }
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(h->stream));
    mark_geometry_dirty(h);
    return 0;
}

int lbm_b200_set_fluid_mask(lbm_b200_t* h, const uint8_t* mask)
{
    GUARD(h);
    return apply_fluid_mask(h, mask, h->z_first, h->nslow(), 0, -1);
}
int lbm_b200_set_fluid_mask_literal(lbm_b200_t* h, const uint8_t* mask)
{
    GUARD(h);
    return apply_fluid_mask(h, mask, h->z_first, h->nslow(), 1, -1);
}
int lbm_b200_set_fluid_mask_global(lbm_b200_t* h, const uint8_t* mask, int literal)
{
    GUARD(h);
    return apply_fluid_mask(h, mask, 1, h->zl_global, literal, -1);
}
int lbm_b200_paint_mask(lbm_b200_t* h, const uint8_t* mask, uint16_t id, int literal)
{
    GUARD(h);
    return apply_fluid_mask(h, mask, 1, h->zl_global, literal, (int) id);
}

int lbm_b200_tag_null_cells(lbm_b200_t* h, int literal, uint64_t* n_tagged)
{
    GUARD(h);
    if (n_tagged) *n_tagged = 0;
    const Layout& g = h->g;
    // which arrays take the tags: the collide field's own layer (literal, or already split -- the reference
    // always goes through Domain::cell()), else the shared maps plus the step count for the read-back
    if (literal || (!h->split && h->null_tagged && ((h->steps - h->null_tagged_at) & 1))) TRY(ensure_split(h));
    GeoLayer& L = h->layer[h->cur];
    CU(cudaMemsetAsync(h->d_counters + 4, 0, sizeof(unsigned int), h->stream));
    dim3 grid((g.xl + 127) / 128, g.yl, g.zl);
    dispatch_q(h->Q, [&](auto Qc) {
        tag_null_kernel<decltype(Qc)::value><<<grid, 128, 0, h->stream>>>(L.kind, g, h->z_first, h->zl_global, h->d_counters + 4);
        return 0;
    });
    h->launches++;
    CU(cudaGetLastError());
    unsigned int n = 0;
    CU(cudaMemcpyAsync(&n, h->d_counters + 4, sizeof n, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    if (n_tagged) *n_tagged = n;
    if (n && !h->split) {
        h->null_tagged = true;
        h->null_tagged_at = h->steps;
    }
    // a tagged cell was not fluid before and has no fluid neighbour: link masks and populations do not
    // change, only what the handler maps report
    return 0;
}

// ---- state ------------------------------------------------------------------
static int transfer_populations(lbm_b200* h, double* host, int layout, int field, bool upload, int z_begin = 0, int z_count = -1)
{
    if (!host) return fail(LBM_B200_EINVAL, "population array is null");
    if (field != LBM_B200_COLLIDE_FIELD && field != LBM_B200_STREAM_FIELD) return fail(LBM_B200_EINVAL, "unknown field %d", field);
    const Layout& g = h->g;
    double* dev = h->f[field == LBM_B200_COLLIDE_FIELD ? h->cur : 1 - h->cur];
    const int Q = h->Q;
    if (z_count < 0) z_count = g.zl + 2 - z_begin;
    if (z_begin < 0 || z_count < 0 || z_begin + z_count > g.zl + 2) return fail(LBM_B200_EINVAL, "plane range outside the slab");
    if (layout == LBM_B200_SOA) {
        if (z_begin != 0 || z_count != g.zl + 2) return fail(LBM_B200_EINVAL, "plane ranges use the AoS layout");
        if (g.swap) return fail(LBM_B200_EINVAL, "y-slabs transfer populations in the AoS layout only");
        const size_t n = h->ncell();
        for (int q = 0; q < Q; ++q) {
            double* d = dev + (size_t) q * g.qstride + X_SHIFT;
            double* s = host + (size_t) q * n;
            const size_t rows = (size_t) (g.yl + 2) * (g.zl + 2);
            if (upload) CU(cudaMemcpy2DAsync(d, g.P * sizeof(double), s, (g.xl + 2) * sizeof(double), (g.xl + 2) * sizeof(double), rows, cudaMemcpyHostToDevice, h->stream));
            else CU(cudaMemcpy2DAsync(s, (g.xl + 2) * sizeof(double), d, g.P * sizeof(double), (g.xl + 2) * sizeof(double), rows, cudaMemcpyDeviceToHost, h->stream));
        }
        CU(cudaStreamSynchronize(h->stream));
        return 0;
    }
    if (layout != LBM_B200_AOS) return fail(LBM_B200_EINVAL, "unknown layout %d", layout);
    // chunks of z planes through a device staging buffer of at most ~256 MB
    const size_t plane_vals = (size_t) (g.xl + 2) * (g.yl + 2) * Q;
    int chunk = (int) std::max<size_t>(1, ((size_t) 256 << 20) / (plane_vals * sizeof(double)));
    chunk = std::min(chunk, std::max(1, z_count));
    DevBuf stage_buf;
    CU(cudaMalloc(&stage_buf.p, plane_vals * chunk * sizeof(double)));
    double* stage = stage_buf.as<double>();
    int rc = 0;
    for (int z0 = z_begin; z0 < z_begin + z_count && rc == 0; z0 += chunk) {
        const int nz = std::min(chunk, z_begin + z_count - z0);
        double* hp = host + plane_vals * (z0 - z_begin);
        cudaError_t e = cudaSuccess;
        if (upload) e = cudaMemcpyAsync(stage, hp, plane_vals * nz * sizeof(double), cudaMemcpyHostToDevice, h->stream);
        if (e == cudaSuccess) {
            dispatch_q(Q, [&](auto Qc) {
                constexpr int QQ = decltype(Qc)::value;
                if (upload) transpose_aos_kernel<QQ, true><<<map_grid(g, nz), 128, 0, h->stream>>>(stage, dev, g, z0);
                else transpose_aos_kernel<QQ, false><<<map_grid(g, nz), 128, 0, h->stream>>>(stage, dev, g, z0);
                return 0;
            });
            h->launches++;
            e = cudaGetLastError();
        }
        if (e == cudaSuccess && !upload) e = cudaMemcpyAsync(hp, stage, plane_vals * nz * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
        if (e != cudaSuccess) rc = fail(LBM_B200_ECUDA, "population transfer failed: %s", cudaGetErrorString(e));
    }
    return rc;
}

int lbm_b200_upload_populations(lbm_b200_t* h, const double* f, int layout, int field)
{
    GUARD(h);
    TRY(transfer_populations(h, const_cast<double*>(f), layout, field, true));
    if (field == LBM_B200_COLLIDE_FIELD) {
        h->first = true;          // boundary cells now hold host-chosen values
        h->materialized = true;
    }
    return 0;
}

// x-y planes [z_begin, z_begin + z_count) of the slab only (AoS): what Domain::cell() needs when a few
// cells of a large lattice are touched
int lbm_b200_upload_planes(lbm_b200_t* h, const double* f, int field, uint64_t z_begin, uint64_t z_count)
{
    GUARD(h);
    if (field == LBM_B200_COLLIDE_FIELD) TRY(materialize(h));   // the other planes' boundary cells must be current
    TRY(transfer_populations(h, const_cast<double*>(f), LBM_B200_AOS, field, true, (int) z_begin, (int) z_count));
    if (field == LBM_B200_COLLIDE_FIELD) {
        h->first = true;
        h->materialized = true;
    }
    return 0;
}
int lbm_b200_download_planes(lbm_b200_t* h, double* f, int field, uint64_t z_begin, uint64_t z_count)
{
    GUARD(h);
    if (field == LBM_B200_COLLIDE_FIELD) TRY(materialize(h));
    return transfer_populations(h, f, LBM_B200_AOS, field, false, (int) z_begin, (int) z_count);
}

int lbm_b200_download_populations(lbm_b200_t* h, double* f, int layout, int field)
{
    GUARD(h);
    if (field == LBM_B200_COLLIDE_FIELD) TRY(materialize(h));
    return transfer_populations(h, f, layout, field, false);
}

int lbm_b200_init_equilibrium(lbm_b200_t* h, const double* rho, const double* u)
{
    GUARD(h);
    if (!rho || !u) return fail(LBM_B200_EINVAL, "rho / u array is null");
    const Layout& g = h->g;
    const size_t plane_cells = (size_t) (g.xl + 2) * (g.yl + 2);
    int chunk = (int) std::max<size_t>(1, ((size_t) 256 << 20) / (plane_cells * 4 * sizeof(double)));
    chunk = std::min(chunk, g.zl + 2);
    DevBuf rho_buf, u_buf;
    CU(cudaMalloc(&rho_buf.p, plane_cells * chunk * sizeof(double)));
    CU(cudaMalloc(&u_buf.p, plane_cells * chunk * 3 * sizeof(double)));
    double *d_rho = rho_buf.as<double>(), *d_u = u_buf.as<double>();
    int rc = 0;
    for (int z0 = 0; z0 < g.zl + 2 && rc == 0; z0 += chunk) {
        const int nz = std::min(chunk, g.zl + 2 - z0);
        cudaError_t e = cudaMemcpyAsync(d_rho, rho + plane_cells * z0, plane_cells * nz * sizeof(double), cudaMemcpyHostToDevice, h->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_u, u + plane_cells * z0 * 3, plane_cells * nz * 3 * sizeof(double), cudaMemcpyHostToDevice, h->stream);
        if (e == cudaSuccess) {
            dispatch_q(h->Q, [&](auto Qc) {
                equilibrium_kernel<decltype(Qc)::value><<<map_grid(g, nz), 128, 0, h->stream>>>(d_rho, d_u, h->f[h->cur], g, z0);
                return 0;
            });
            h->launches++;
            e = cudaGetLastError();
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
        if (e != cudaSuccess) rc = fail(LBM_B200_ECUDA, "equilibrium initialisation failed: %s", cudaGetErrorString(e));
    }
    if (rc == 0) { h->first = true; h->materialized = true; }
    return rc;
}

// ---- checkpoint / restart (absent in the reference) ------------------------------
// File: 64-byte header {magic "LBMB200\0", version, Q, xl, yl, zl_local, z_first, zl_global, arithmetic mode,
// steps, tau, checksum of the collide field's handler-kind map} followed by the collide field and then the
// stream field, each as Q dense arrays in Domain::idx order (layout LBM_B200_SOA).  The stream field carries
// state too: ghost-shell cells that kept the fluid handler are collided in place in both lattices.
namespace {
struct CheckpointHeader {
    char magic[8];
    int32_t version, Q;
    int32_t xl, yl, zl_local, z_first, zl_global, arithmetic;
    uint64_t steps;
    double tau;
    uint64_t kind_hash;
};
static_assert(sizeof(CheckpointHeader) == 64, "checkpoint header layout");

// FNV-1a over the dense kind map of one layer (tags of set_nonfluid_cells_nullcollide read as the former kind,
// so that the checksum does not depend on the step parity)
int hash_layer(lbm_b200* h, const GeoLayer& L, uint64_t* out)
{
    const Layout& g = h->g;
    const size_t n = h->ncell();
    DevBuf b8;
    CU(cudaMalloc(&b8.p, n));
    gather_maps_kernel<<<map_grid(g, g.zl + 2), 128, 0, h->stream>>>(L.kind, L.bcid, h->d_bc, (int) h->h_bc.size(), b8.as<uint8_t>(),
                                                                    nullptr, g, 0, h->split ? 0 : 1);
    h->launches++;
    CU(cudaGetLastError());
    std::vector<uint8_t> k(n);
    CU(cudaMemcpyAsync(k.data(), b8.p, n, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    uint64_t x = 1469598103934665603ull;
    for (uint8_t b : k) { x ^= b; x *= 1099511628211ull; }
    *out = x;
    return 0;
}
// both lattices, independent of which one is the collide field at the moment
int kind_hash(lbm_b200* h, uint64_t* out)
{
    uint64_t a = 0, b = 0;
    TRY(hash_layer(h, h->layer[0], &a));
    if (h->split) TRY(hash_layer(h, h->layer[1], &b));
    *out = a + b;
    return 0;
}
}

int lbm_b200_save_checkpoint(lbm_b200_t* h, const char* path)
{
    GUARD(h);
    if (!path) return fail(LBM_B200_EINVAL, "null path");
    TRY(materialize(h));
    const Layout& g = h->g;
    CheckpointHeader hd{};
    memcpy(hd.magic, "LBMB200", 8);
    hd.version = 2 + (h->axis == 1 ? 0x100 : 0); hd.Q = h->Q;       // y-slabs store x-z planes: another file layout
    hd.xl = g.xl; hd.yl = g.yl; hd.zl_local = g.zl; hd.z_first = h->z_first; hd.zl_global = h->zl_global;
    hd.arithmetic = h->exact ? LBM_B200_EXACT : LBM_B200_FAST;
    hd.steps = h->steps; hd.tau = h->tau;
    TRY(kind_hash(h, &hd.kind_hash));
    FILE* fp = fopen(path, "wb");
    if (!fp) return fail(LBM_B200_EINVAL, "cannot open %s for writing", path);
    int rc = fwrite(&hd, sizeof hd, 1, fp) == 1 ? 0 : fail(LBM_B200_EINVAL, "write to %s failed", path);
    const size_t n = h->ncell();
    std::vector<double> buf(n);
    const size_t rows = (size_t) (g.yl + 2) * (g.zl + 2);
    for (int field = 0; field < 2 && rc == 0; ++field) {
        const double* base = h->f[field == 0 ? h->cur : 1 - h->cur];
        for (int q = 0; q < h->Q && rc == 0; ++q) {
            const double* d = base + (size_t) q * g.qstride + X_SHIFT;
            cudaError_t e = cudaMemcpy2DAsync(buf.data(), (g.xl + 2) * sizeof(double), d, g.P * sizeof(double),
                                              (g.xl + 2) * sizeof(double), rows, cudaMemcpyDeviceToHost, h->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
            if (e != cudaSuccess) rc = fail(LBM_B200_ECUDA, "checkpoint download failed: %s", cudaGetErrorString(e));
            else if (fwrite(buf.data(), sizeof(double), n, fp) != n) rc = fail(LBM_B200_EINVAL, "write to %s failed", path);
        }
    }
    if (fclose(fp) != 0 && rc == 0) rc = fail(LBM_B200_EINVAL, "closing %s failed", path);
    return rc;
}

int lbm_b200_load_checkpoint(lbm_b200_t* h, const char* path)
{
    GUARD(h);
    if (!path) return fail(LBM_B200_EINVAL, "null path");
    const Layout& g = h->g;
    FILE* fp = fopen(path, "rb");
    if (!fp) return fail(LBM_B200_EINVAL, "cannot open %s", path);
    CheckpointHeader hd{};
    int rc = 0;
    uint64_t my_hash = 0;
    if (fread(&hd, sizeof hd, 1, fp) != 1 || memcmp(hd.magic, "LBMB200", 8) != 0 || (hd.version & 0xff) != 2)
        rc = fail(LBM_B200_EINVAL, "%s is not a lbm_b200 checkpoint (version 2)", path);
    else if (hd.version != 2 + (h->axis == 1 ? 0x100 : 0))
        rc = fail(LBM_B200_EINVAL, "%s was written by a slab with another split axis", path);
    else if (hd.Q != h->Q || hd.xl != g.xl || hd.yl != g.yl || hd.zl_local != g.zl || hd.z_first != h->z_first || hd.zl_global != h->zl_global)
        rc = fail(LBM_B200_EINVAL, "%s holds D3Q%d %dx%dx%d (slab at %d of %d), this domain is D3Q%d %dx%dx%d (slab at %d of %d)", path,
                  hd.Q, hd.xl, hd.yl, hd.zl_local, hd.z_first, hd.zl_global, h->Q, g.xl, g.yl, g.zl, h->z_first, h->zl_global);
    else if (hd.arithmetic != (h->exact ? LBM_B200_EXACT : LBM_B200_FAST))
        rc = fail(LBM_B200_EINVAL, "%s was written in %s arithmetic, this domain runs in %s arithmetic", path,
                  hd.arithmetic == LBM_B200_EXACT ? "exact" : "fast", h->exact ? "exact" : "fast");
    else if ((rc = kind_hash(h, &my_hash)) == 0 && my_hash != hd.kind_hash)
        rc = fail(LBM_B200_EINVAL, "%s was written under a different geometry (handler-map checksum differs): re-apply the scenario first", path);
    const size_t n = h->ncell();
    std::vector<double> buf(rc == 0 ? n : 0);
    const size_t rows = (size_t) (g.yl + 2) * (g.zl + 2);
    for (int field = 0; field < 2 && rc == 0; ++field) {
        double* base = h->f[field == 0 ? h->cur : 1 - h->cur];
        for (int q = 0; q < h->Q && rc == 0; ++q) {
            if (fread(buf.data(), sizeof(double), n, fp) != n) { rc = fail(LBM_B200_EINVAL, "%s is truncated", path); break; }
            double* d = base + (size_t) q * g.qstride + X_SHIFT;
            cudaError_t e = cudaMemcpy2DAsync(d, g.P * sizeof(double), buf.data(), (g.xl + 2) * sizeof(double),
                                              (g.xl + 2) * sizeof(double), rows, cudaMemcpyHostToDevice, h->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
            if (e != cudaSuccess) rc = fail(LBM_B200_ECUDA, "checkpoint upload failed: %s", cudaGetErrorString(e));
        }
    }
    fclose(fp);
    if (rc == 0) {
        // handlers belong to lattices: if the file's collide field is the other lattice (odd distance in steps),
        // the layers trade places; null tags follow the absolute step count by themselves
        if (h->split && ((hd.steps - h->steps) & 1)) {
            std::swap(h->layer[0], h->layer[1]);
            mark_geometry_dirty(h);
        }
        h->steps = hd.steps;
        h->first = true;          // boundary cells hold the stored (materialised) values again
        h->materialized = true;
    }
    return rc;
}

// ---- hot path ---------------------------------------------------------------
int lbm_b200_step(lbm_b200_t* h, uint64_t n_steps)
{
    GUARD(h);
    TRY(ready_to_step(h));
    CU(cudaEventRecord(h->ev_a, h->stream));
    TRY(enqueue_steps(h, n_steps));
    CU(cudaEventRecord(h->ev_b, h->stream));
    h->timed = true;
    return 0;
}

int lbm_b200_step_group(lbm_b200_t* const* hs, int n, uint64_t n_steps)
{
    if (n < 0 || (n > 0 && !hs)) return fail(LBM_B200_EINVAL, "bad handle list");
    for (int i = 0; i < n; ++i)
        if (!hs[i]) return fail(LBM_B200_EINVAL, "null handle");
    int prev = -1;
    cudaGetDevice(&prev);
    int rc = 0;
    for (int i = 0; i < n && rc == 0; ++i) {
        if (cudaSetDevice(hs[i]->device) != cudaSuccess) rc = fail(LBM_B200_ECUDA, "cannot select CUDA device %d", hs[i]->device);
        if (rc == 0) rc = ready_to_step(hs[i]);
        if (rc == 0 && cudaEventRecord(hs[i]->ev_a, hs[i]->stream) != cudaSuccess) rc = fail(LBM_B200_ECUDA, "cudaEventRecord failed");
    }
    // Runs of at most GRAPH_STEPS steps per slab, slab after slab: a slab's wait kernel never sits in front
    // of a device queue for longer than it takes the host to submit the same run to its neighbours.
    uint64_t done = 0;
    while (rc == 0 && done < n_steps) {
        const uint64_t run = std::min<uint64_t>(lbm_b200::GRAPH_STEPS, n_steps - done);
        for (int i = 0; i < n && rc == 0; ++i) {
            if (cudaSetDevice(hs[i]->device) != cudaSuccess) rc = fail(LBM_B200_ECUDA, "cannot select CUDA device %d", hs[i]->device);
            if (rc == 0) rc = enqueue_steps(hs[i], run);
        }
        done += run;
    }
    for (int i = 0; i < n && rc == 0; ++i) {
        if (cudaSetDevice(hs[i]->device) != cudaSuccess || cudaEventRecord(hs[i]->ev_b, hs[i]->stream) != cudaSuccess)
            rc = fail(LBM_B200_ECUDA, "cudaEventRecord failed");
        hs[i]->timed = true;
    }
    if (prev >= 0) cudaSetDevice(prev);
    return rc;
}

int lbm_b200_sync(lbm_b200_t* h)
{
    GUARD(h);
    CU(cudaStreamSynchronize(h->stream));
    return halo_check(h);
}

int lbm_b200_elapsed_ms(lbm_b200_t* h, double* ms)
{
    GUARD(h);
    if (!ms) return fail(LBM_B200_EINVAL, "null output pointer");
    if (!h->timed) return fail(LBM_B200_ESTATE, "no step has been timed yet");
    CU(cudaEventSynchronize(h->ev_b));
    float t = 0.f;
    CU(cudaEventElapsedTime(&t, h->ev_a, h->ev_b));
    *ms = t;
    return 0;
}

int lbm_b200_launch_count(lbm_b200_t* h, uint64_t* n)
{
    if (!h || !n) return fail(LBM_B200_EINVAL, "null argument");
    *n = h->launches;
    return 0;
}
int lbm_b200_tma_launch_count(lbm_b200_t* h, uint64_t* n)
{
    if (!h || !n) return fail(LBM_B200_EINVAL, "null argument");
    *n = h->tma_launches;
    return 0;
}
uint64_t lbm_b200_steps_done(lbm_b200_t* h) { return h ? h->steps : 0; }

// ---- read-out -----------------------------------------------------------------
int lbm_b200_macroscopic_end(lbm_b200_t* h)
{
    GUARD(h);
    if (!h->readout_pending) return 0;
    h->readout_pending = false;
    CU(cudaEventSynchronize(h->ev_copied));
    return 0;
}

int lbm_b200_macroscopic_begin(lbm_b200_t* h, double* rho, double* u)
{
    GUARD(h);
    TRY(lbm_b200_macroscopic_end(h));     // the staging arrays are about to be overwritten
    TRY(materialize(h));
    const Layout& g = h->g;
    const size_t n = (size_t) g.xl * g.yl * g.zl;
    if (rho && !h->d_rho) CU(cudaMalloc(&h->d_rho, n * sizeof(double)));
    if (u && !h->d_u) CU(cudaMalloc(&h->d_u, 3 * n * sizeof(double)));
    double* d_rho = rho ? h->d_rho : nullptr;
    double* d_u = u ? h->d_u : nullptr;
    // z chunks: the reduction of chunk k+1 (compute stream) runs while chunk k crosses PCIe (copy stream);
    // the whole snapshot is reduced before any later step touches the lattice, the copies then overlap
    // with those steps
    const int chunks = std::min(lbm_b200::READOUT_CHUNKS, g.zl);
    const size_t plane = (size_t) g.xl * g.yl;
    for (int c = 0; c < chunks; ++c) {
        const int z0 = (int) ((long long) g.zl * c / chunks), z1 = (int) ((long long) g.zl * (c + 1) / chunks);
        if (z1 <= z0) continue;
        dim3 grid((g.xl + 127) / 128, g.yl, z1 - z0);
        dispatch_q(h->Q, [&](auto Qc) {
            macroscopic_kernel<decltype(Qc)::value><<<grid, 128, 0, h->stream>>>(h->f[h->cur], g, d_rho, d_u, z0);
            return 0;
        });
        h->launches++;
        CU(cudaGetLastError());
        CU(cudaEventRecord(h->ev_chunk[c], h->stream));
        // the 4 doubles per cell of this chunk (1 density + 3 velocity) are dealt to the copy streams in equal pieces
        const size_t off = plane * z0, cnt = plane * (z1 - z0);
        const int ns = h->n_copy_streams;
        auto stream_of = [&](int k) { return k == 0 ? h->copy_stream : h->copy_extra[k - 1]; };
        for (int k = 0; k < ns; ++k) CU(cudaStreamWaitEvent(stream_of(k), h->ev_chunk[c], 0));
        // pieces: density = piece 0, velocity split into 3 (component-interleaved, so split by cells)
        int piece = 0;
        if (rho) CU(cudaMemcpyAsync(rho + off, d_rho + off, cnt * sizeof(double), cudaMemcpyDeviceToHost, stream_of(piece++ % ns)));
        if (u) {
            const int parts = 3;
            for (int k = 0; k < parts; ++k) {
                const size_t a = cnt * k / parts, b = cnt * (k + 1) / parts;
                if (b > a) CU(cudaMemcpyAsync(u + 3 * (off + a), d_u + 3 * (off + a), 3 * (b - a) * sizeof(double), cudaMemcpyDeviceToHost,
                                              stream_of(piece++ % ns)));
            }
        }
    }
    for (int k = 0; k + 1 < h->n_copy_streams; ++k) {
        CU(cudaEventRecord(h->ev_extra[k], h->copy_extra[k]));
        CU(cudaStreamWaitEvent(h->copy_stream, h->ev_extra[k], 0));
    }
    CU(cudaEventRecord(h->ev_copied, h->copy_stream));
    h->readout_pending = true;
    return 0;
}

int lbm_b200_macroscopic(lbm_b200_t* h, double* rho, double* u)
{
    TRY(lbm_b200_macroscopic_begin(h, rho, u));
    return lbm_b200_macroscopic_end(h);
}

// ---- host memory next to a GPU ---------------------------------------------------------------------
namespace {
std::mutex g_host_mutex;
std::map<void*, size_t> g_host_blocks;     // pointer -> bytes (posix_memalign + cudaHostRegister)

// CPUs of the NUMA node the device hangs off; false where the topology is unknown (containers, 1 node)
bool device_cpus(int device, cpu_set_t* set)
{
    char bus[32] = {};
    if (cudaDeviceGetPCIBusId(bus, sizeof bus, device) != cudaSuccess) { cudaGetLastError(); return false; }
    for (char* c = bus; *c; ++c) *c = (char) tolower(*c);
    char path[128];
    snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/numa_node", bus);
    FILE* fp = fopen(path, "r");
    if (!fp) return false;
    int node = -1;
    const int got = fscanf(fp, "%d", &node);
    fclose(fp);
    if (got != 1 || node < 0) return false;
    snprintf(path, sizeof path, "/sys/devices/system/node/node%d/cpulist", node);
    fp = fopen(path, "r");
    if (!fp) return false;
    char list[4096] = {};
    const bool ok = fgets(list, sizeof list, fp) != nullptr;
    fclose(fp);
    if (!ok) return false;
    CPU_ZERO(set);
    int count = 0;
    for (char* tok = strtok(list, ",\n"); tok; tok = strtok(nullptr, ",\n")) {
        int a = 0, b = 0;
        const int k = sscanf(tok, "%d-%d", &a, &b);
        if (k == 1) b = a;
        if (k < 1) continue;
        for (int c = a; c <= b && c < CPU_SETSIZE; ++c) { CPU_SET(c, set); ++count; }
    }
    return count > 0;
}
}

int lbm_b200_selftest_division(uint64_t n, uint64_t seed, double tau, uint64_t* mismatches)
{
    if (!mismatches || !(tau > 0.0)) return fail(LBM_B200_EINVAL, "bad self-test request");
    unsigned long long* d = nullptr;
    if (cudaMalloc(&d, sizeof *d) != cudaSuccess) return fail(LBM_B200_ECUDA, "no CUDA device for the division self-test");
    cudaMemset(d, 0, sizeof *d);
    for (int which = 0; which < 5; ++which)
        division_selftest_kernel<<<148 * 8, 256>>>(n, seed + which, which, tau, 1.0 / tau, d);
    unsigned long long bad = 0;
    const cudaError_t e = cudaMemcpy(&bad, d, sizeof bad, cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) return fail(LBM_B200_ECUDA, "division self-test failed: %s", cudaGetErrorString(e));
    *mismatches = bad;
    return 0;
}

int lbm_b200_bind_host_thread(int device)
{
    if (device < 0 && cudaGetDevice(&device) != cudaSuccess) return fail(LBM_B200_ECUDA, "cudaGetDevice failed");
    cpu_set_t set;
    if (device_cpus(device, &set)) sched_setaffinity(0, sizeof set, &set);
    return 0;
}

int lbm_b200_host_alloc(void** ptr, size_t bytes, int device)
{
    if (!ptr || bytes == 0) return fail(LBM_B200_EINVAL, "bad host allocation request");
    *ptr = nullptr;
    if (device < 0 && cudaGetDevice(&device) != cudaSuccess) return fail(LBM_B200_ECUDA, "cudaGetDevice failed");
    // first touch from a thread pinned to the GPU's node places the pages there; then page-lock them
    cpu_set_t old_set, near_set;
    const bool have_old = sched_getaffinity(0, sizeof old_set, &old_set) == 0;
    const bool moved = have_old && device_cpus(device, &near_set) && sched_setaffinity(0, sizeof near_set, &near_set) == 0;
    void* p = nullptr;
    const size_t rounded = (bytes + 4095) / 4096 * 4096;
    int rc = 0;
    if (posix_memalign(&p, 4096, rounded) != 0) rc = fail(LBM_B200_ENOMEM, "host allocation of %zu bytes failed", bytes);
    if (rc == 0) {
        for (size_t o = 0; o < rounded; o += 4096) ((volatile char*) p)[o] = 0;
        const cudaError_t e = cudaHostRegister(p, rounded, cudaHostRegisterPortable);
        if (e != cudaSuccess) {
            free(p);
            cudaGetLastError();
            rc = fail(LBM_B200_ECUDA, "cudaHostRegister of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
        }
    }
    if (moved) sched_setaffinity(0, sizeof old_set, &old_set);
    if (rc != 0) return rc;
    std::lock_guard<std::mutex> lock(g_host_mutex);
    g_host_blocks[p] = rounded;
    *ptr = p;
    return 0;
}

int lbm_b200_host_free(void* ptr)
{
    if (!ptr) return 0;
    std::lock_guard<std::mutex> lock(g_host_mutex);
    auto it = g_host_blocks.find(ptr);
    if (it == g_host_blocks.end()) return fail(LBM_B200_EINVAL, "pointer was not returned by lbm_b200_host_alloc");
    cudaHostUnregister(ptr);
    cudaGetLastError();
    free(ptr);
    g_host_blocks.erase(it);
    return 0;
}

int lbm_b200_diagnostics(lbm_b200_t* h, double* mass, double* kinetic, double* umax)
{
    GUARD(h);
    TRY(commit_geometry(h));
    const Layout& g = h->g;
    dim3 grid((g.xl + 127) / 128, g.yl, g.zl);
    const size_t nblk = (size_t) grid.x * grid.y * grid.z;
    DevBuf part_buf;
    CU(cudaMalloc(&part_buf.p, nblk * 3 * sizeof(double)));
    double* d_part = part_buf.as<double>();
    dispatch_q(h->Q, [&](auto Qc) {
        diagnostics_kernel<decltype(Qc)::value><<<grid, 128, 0, h->stream>>>(h->f[h->cur], h->layer[h->cur].kind, g, d_part);
        return 0;
    });
    h->launches++;
    std::vector<double> part(nblk * 3);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(part.data(), d_part, nblk * 3 * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    if (e != cudaSuccess) return fail(LBM_B200_ECUDA, "diagnostics failed: %s", cudaGetErrorString(e));
    double m = 0.0, k = 0.0, um = 0.0;
    for (size_t b = 0; b < nblk; ++b) { m += part[3 * b]; k += part[3 * b + 1]; um = std::max(um, part[3 * b + 2]); }
    if (mass) *mass = m;
    if (kinetic) *kinetic = k;
    if (umax) *umax = std::sqrt(um);
    return 0;
}

// ---- multi-GPU z-slabs ----------------------------------------------------------
int lbm_b200_halo_layout(lbm_b200_t* h, int* n_q, size_t* bytes)
{
    if (!h) return fail(LBM_B200_EINVAL, "null handle");
    const int n = dispatch_q(h->Q, [&](auto Qc) { return Lattice<decltype(Qc)::value>::n_up(); });
    if (n_q) *n_q = n;
    if (bytes) *bytes = (size_t) h->g.plane * sizeof(double);
    return 0;
}

int lbm_b200_halo_plane(lbm_b200_t* h, int buffer, int side, int k, int recv, void** ptr)
{
    if (!h || !ptr) return fail(LBM_B200_EINVAL, "null argument");
    if (buffer < 0 || buffer > 1 || side < 0 || side > 1) return fail(LBM_B200_EINVAL, "bad buffer/side");
    // the k-th population with c_z = +1 (side UP sends those; side DOWN receives them) or -1
    const int want_send = side == LBM_B200_UP ? 1 : -1;
    const int want = recv ? -want_send : want_send;
    int q_found = -1;
    dispatch_q(h->Q, [&](auto Qc) {
        using L = Lattice<decltype(Qc)::value>;
        int c = 0;
        for (int q = 0; q < h->Q; ++q)
            if ((h->axis == 1 ? L::cy(q) : L::cz(q)) == want) { if (c == k) q_found = q; ++c; }
        return 0;
    });
    if (q_found < 0) return fail(LBM_B200_EINVAL, "halo population index %d out of range", k);
    const Layout& g = h->g;
    int z;
    if (side == LBM_B200_UP) z = recv ? h->nslow() + 1 : h->nslow();
    else z = recv ? 0 : 1;
    // the cells of plane z occupy [z*plane + X_SHIFT, (z+1)*plane + X_SHIFT): rows are shifted by X_SHIFT elements
    *ptr = h->f[buffer] + (size_t) q_found * g.qstride + (size_t) z * g.plane + X_SHIFT;
    return 0;
}

int lbm_b200_dst_buffer(lbm_b200_t* h) { return h ? 1 - h->cur : -1; }

int lbm_b200_edge_plane(lbm_b200_t* h, int side, int q, int recv, void** ptr)
{
    if (!h || !ptr) return fail(LBM_B200_EINVAL, "null argument");
    if (side < 0 || side > 1 || q < 0 || q >= h->Q) return fail(LBM_B200_EINVAL, "bad side / population");
    const Layout& g = h->g;
    const int z = side == LBM_B200_UP ? (recv ? h->nslow() + 1 : h->nslow()) : (recv ? 0 : 1);
    *ptr = h->f[h->cur] + (size_t) q * g.qstride + (size_t) z * g.plane + X_SHIFT;
    return 0;
}

int lbm_b200_step_edges(lbm_b200_t* h)
{
    GUARD(h);
    if (h->edges_done) return fail(LBM_B200_ESTATE, "step_edges called twice");
    TRY(ready_to_step(h));
    if (h->nslow() > 1) TRY(launch_sweep(h, 1, 2, true, h->nslow() - 1, interior_mode(h)));
    else TRY(launch_sweep(h, 1, 1, true, 1, interior_mode(h)));
    h->edges_done = true;
    return 0;
}
int lbm_b200_step_interior(lbm_b200_t* h)
{
    GUARD(h);
    if (!h->edges_done) return fail(LBM_B200_ESTATE, "step_interior before step_edges");
    TRY(sweep_planes(h, 2, h->nslow() - 2, true));
    TRY(launch_inplace(h));
    return 0;
}
int lbm_b200_step_finish(lbm_b200_t* h)
{
    if (!h) return fail(LBM_B200_EINVAL, "null handle");
    if (!h->edges_done) return fail(LBM_B200_ESTATE, "step_finish before step_edges");
    h->edges_done = false;
    finish_step(h);
    return 0;
}

// Read-back support for multi-slab domains: copy ALL Q populations of this slab's two edge planes
// (current collide field) into the neighbours' ghost planes, so that their boundary pass can treat
// cells next to the cut like the reference does.  The caller synchronises every slab (lbm_b200_sync
// + a barrier across processes) between this call and the read-back, and calls it on every slab.
int lbm_b200_halo_push_all(lbm_b200_t* h)
{
    GUARD(h);
    const Layout& g = h->g;
    const size_t bytes = (size_t) g.plane * sizeof(double);
    for (int side = 0; side < 2; ++side) {
        double* peer = h->peer_f[side][h->cur];
        if (!peer) continue;
        const int z_mine = side == LBM_B200_UP ? h->nslow() : 1;
        for (int q = 0; q < h->Q; ++q)
            CU(cudaMemcpyAsync(peer + (size_t) q * h->peer_qstride[side] + h->peer_off[side],
                               h->f[h->cur] + (size_t) q * g.qstride + (size_t) z_mine * g.plane, bytes,
                               cudaMemcpyDeviceToDevice, h->stream));
    }
    return 0;
}
// tells this slab that its neighbours have pushed (and the caller has synchronised)
int lbm_b200_halo_pushed(lbm_b200_t* h)
{
    if (!h) return fail(LBM_B200_EINVAL, "null handle");
    h->full_halo_at = h->steps;
    if (!h->first) h->materialized = false;     // (while `first` is set the boundary cells hold host-visible values: nothing to redo)
    return 0;
}

int lbm_b200_export(lbm_b200_t* h, void* blob)
{
    GUARD(h);
    if (!blob) return fail(LBM_B200_EINVAL, "null blob");
    memset(blob, 0, LBM_B200_EXPORT_BYTES);
    unsigned char* p = (unsigned char*) blob;
    cudaIpcMemHandle_t mh;
    CU(cudaIpcGetMemHandle(&mh, h->f[0]));
    static_assert(sizeof(mh) == 64, "CUDA IPC handle size");
    memcpy(p, &mh, 64);
    long long meta[4] = { h->g.qstride, h->g.plane, h->nslow() + 1000000LL * h->axis, h->Q };
    memcpy(p + 64, meta, sizeof meta);
    return 0;
}

static int connect_common(lbm_b200* h, int side, double* base, unsigned long long* nb_flags, long long qstride,
                          long long plane, long long zl, long long Q)
{
    const long long nb_axis = zl / 1000000LL;      // (the split axis travels in the upper digits of the plane count)
    zl %= 1000000LL;
    if (Q != h->Q || plane != h->g.plane || nb_axis != h->axis)
        return fail(LBM_B200_EINVAL, "neighbour slab has a different lattice, plane shape or split axis");
    // we are the neighbour's DOWN side when it is our UP side, and vice versa
    h->peer_flag[side] = nb_flags + (side == LBM_B200_UP ? LBM_B200_DOWN : LBM_B200_UP);
    h->peer_f[side][0] = base;
    h->peer_f[side][1] = base + (size_t) qstride * Q;
    h->peer_qstride[side] = qstride;
    // my top plane feeds the upper neighbour's ghost plane 0; my bottom plane the
    // lower neighbour's ghost plane zl_nb+1
    h->peer_off[side] = side == LBM_B200_UP ? 0 : (zl + 1) * plane;
    drop_graphs(h);
    return 0;
}

int lbm_b200_connect(lbm_b200_t* h, int side, const void* blob)
{
    GUARD(h);
    if (side < 0 || side > 1 || !blob) return fail(LBM_B200_EINVAL, "bad side / blob");
    const unsigned char* p = (const unsigned char*) blob;
    cudaIpcMemHandle_t mh;
    memcpy(&mh, p, 64);
    long long meta[4];
    memcpy(meta, p + 64, sizeof meta);
    void* base = nullptr;
    const int other = 1 - side;
    if (h->peer_ipc_base[other] && memcmp(&h->peer_ipc_handle[other], &mh, sizeof mh) == 0) {
        base = h->peer_ipc_base[other];          // ring of two: the same neighbour on both sides, map it once
        h->peer_ipc_shared = true;
    } else {
        CU(cudaIpcOpenMemHandle(&base, mh, cudaIpcMemLazyEnablePeerAccess));
    }
    h->peer_ipc_base[side] = base;
    h->peer_ipc_handle[side] = mh;
    unsigned long long* nb_flags = reinterpret_cast<unsigned long long*>((double*) base + 2 * meta[0] * meta[3]);
    return connect_common(h, side, (double*) base, nb_flags, meta[0], meta[1], meta[2], meta[3]);
}

// Drops the peer mappings (after the caller has synchronised every slab): a slab must not be
// destroyed while a neighbour can still store into it.
int lbm_b200_disconnect(lbm_b200_t* h)
{
    GUARD(h);
    CU(cudaStreamSynchronize(h->stream));
    drop_graphs(h);
    if (h->peer_ipc_shared) h->peer_ipc_base[1] = nullptr;   // mapped once
    h->peer_ipc_shared = false;
    for (int s = 0; s < 2; ++s) {
        if (h->peer_ipc_base[s]) cudaIpcCloseMemHandle(h->peer_ipc_base[s]);
        h->peer_ipc_base[s] = nullptr;
        h->peer_f[s][0] = h->peer_f[s][1] = nullptr;
        h->peer_flag[s] = nullptr;
    }
    cudaGetLastError();
    return 0;
}

int lbm_b200_connect_local(lbm_b200_t* h, int side, lbm_b200_t* nb)
{
    GUARD(h);
    if (side < 0 || side > 1 || !nb) return fail(LBM_B200_EINVAL, "bad side / neighbour");
    if (nb->device != h->device) {
        int can = 0;
        CU(cudaDeviceCanAccessPeer(&can, h->device, nb->device));
        if (!can) return fail(LBM_B200_ECUDA, "device %d cannot access device %d", h->device, nb->device);
        cudaError_t e = cudaDeviceEnablePeerAccess(nb->device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(LBM_B200_ECUDA, "cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e));
        cudaGetLastError();
    }
    return connect_common(h, side, nb->f[0], nb->d_flags, nb->g.qstride, nb->g.plane, nb->nslow() + 1000000LL * nb->axis, nb->Q);
}

} // extern "C"

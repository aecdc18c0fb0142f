// engine.cu -- host side of the C ABI declared in include/lbm_b200.h.
//
// One handle = one Domain (domain.h:12-19) or one z-slab of it, resident on one
// B200.  The reference's stream(); swap(); collide(); (src/main.cpp:50-52) is a
// single launch of sweep_kernel per step plus a buffer-index flip.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>
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

struct lbm_b200 {
    int Q = 0;
    int device = 0;
    Layout g{};
    int zl_global = 0;
    int z_first = 1;           // global index of local plane 1
    double tau = 1.0;
    int exact = 0;

    double* f[2] = { nullptr, nullptr };
    int cur = 0;               // f[cur] is the collide field
    uint32_t* d_mask = nullptr;
    uint32_t* d_bits = nullptr;
    uint8_t* d_kind = nullptr;
    uint16_t* d_bcid = nullptr;
    BcRec* d_bc = nullptr;
    int* d_ghost = nullptr;
    int n_ghost = 0;

    std::vector<uint8_t> h_kind;     // dense, local idx order
    std::vector<uint16_t> h_bcid;
    std::vector<lbm_b200_bc> h_bc;
    bool geom_dirty = true;
    bool geom_unchecked = false;   // maps supplied as arrays, not validated yet
    uint8_t* stage_kind = nullptr; // dense copies already on the device (set_geometry), consumed by commit
    uint16_t* stage_bcid = nullptr;
    double* d_rho = nullptr;       // persistent read-out staging
    double* d_u = nullptr;
    bool first = true;         // boundary cells hold host-visible (stored) values
    bool materialized = true;  // boundary cells of f[cur] hold the reference's values
    int wrap_z = 0;
    bool ring_lo = false, ring_hi = false;   // periodic z must be closed by a slab ring on that side

    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_a = nullptr, ev_b = nullptr;
    bool timed = false;
    uint64_t steps = 0;
    uint64_t full_halo_at = ~0ull;   // time level for which neighbours pushed their full edge planes
    uint64_t launches = 0;
    bool edges_done = false;   // split-phase state

    // direct peer stores: [side]
    double* peer_f[2][2] = { { nullptr, nullptr }, { nullptr, nullptr } };   // [side][buffer]
    long long peer_qstride[2] = { 0, 0 };
    long long peer_off[2] = { 0, 0 };
    void* peer_ipc_base[2] = { nullptr, nullptr };
    cudaIpcMemHandle_t peer_ipc_handle[2] = {};
    bool peer_ipc_shared = false;                   // both sides map the same exporter (ring of two)
    unsigned long long* d_flags = nullptr;          // [side]: sweeps completed by the neighbour on that side
    unsigned long long* peer_flag[2] = { nullptr, nullptr };   // the neighbour's counter for us
    unsigned long long halo_epoch = 0;              // sweeps completed since the peers were connected
    int* d_halo_error = nullptr;
    int clock_khz = 1965000;
    long long pull_offset[27] = {};                 // c_z*plane + c_y*P + c_x per direction
    unsigned long long* d_trace = nullptr;         // LBM_B200_HALO_TRACE=<file prefix>: wait-kernel timestamps
    static constexpr int TRACE_EPOCHS = 8192;
    int* h_halo_error = nullptr;                    // pinned mirror

    size_t ncell() const { return (size_t) (g.xl + 2) * (g.yl + 2) * (g.zl + 2); }
    size_t field_bytes() const { return (size_t) g.qstride * Q * sizeof(double); }
    size_t map_elems() const { return (size_t) g.qstride; }
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

// Checks the dense host maps (optionally copying them from caller arrays in the same parallel pass):
// known kinds, bc ids inside the table and of the same kind, PERIODIC only on the ghost shell.
int validate_maps(lbm_b200* h, const uint8_t* kind_src, const uint16_t* bcid_src)
{
    const Layout& g = h->g;
    const int nb = (int) h->h_bc.size();
    const size_t rows = (size_t) (g.zl + 2) * (g.yl + 2);
    int bad = 0;
    long long bad_at = -1;
    #pragma omp parallel for schedule(static)
    for (long long r = 0; r < (long long) rows; ++r) {
        const int z = (int) (r / (g.yl + 2)), y = (int) (r % (g.yl + 2));
        const size_t row = (size_t) r * (g.xl + 2);
        if (kind_src) memcpy(&h->h_kind[row], kind_src + row, (size_t) g.xl + 2);
        if (kind_src && bcid_src) memcpy(&h->h_bcid[row], bcid_src + row, ((size_t) g.xl + 2) * sizeof(uint16_t));
        if (kind_src && !bcid_src) memset(&h->h_bcid[row], 0, ((size_t) g.xl + 2) * sizeof(uint16_t));
        for (int x = 0; x < g.xl + 2; ++x) {
            const int k = h->h_kind[row + x];
            int why = 0;
            if (k >= K_COUNT) why = 1;
            else if (k >= K_NOSLIP && k <= K_PRESSURE) {
                const int id = h->h_bcid[row + x];
                if (id >= nb) why = 2;
                else if (h->h_bc[id].kind != k) why = 3;
            } else if (k == K_PERIODIC && x > 0 && x < g.xl + 1 && y > 0 && y < g.yl + 1 && z > 0 && z < g.zl + 1) why = 4;
            if (why) {
                #pragma omp critical
                if (!bad) { bad = why; bad_at = (long long) (row + x); }
            }
        }
    }
    if (bad) {
        const long long x = bad_at % (g.xl + 2), y = (bad_at / (g.xl + 2)) % (g.yl + 2), z = bad_at / ((long long) (g.xl + 2) * (g.yl + 2));
        const char* msg[] = { "", "unknown kind", "bc id outside the table", "kind differs from table[bc id].kind", "PERIODIC is a ghost-shell kind" };
        return fail(LBM_B200_EINVAL, "cell (%lld,%lld,%lld): %s", x, y, z, msg[bad]);
    }
    return 0;
}

// upload maps, boundary table, link mask, ghost-fluid list
int commit_geometry(lbm_b200* h)
{
    if (!h->geom_dirty) return 0;
    const Layout& g = h->g;
    const size_t n = h->ncell();
    // boundary records
    {
        std::vector<BcRec> recs(std::max<size_t>(1, h->h_bc.size()));
        memset(recs.data(), 0, recs.size() * sizeof(BcRec));
        for (size_t i = 0; i < h->h_bc.size(); ++i) {
            recs[i].kind = h->h_bc[i].kind;
            for (int d = 0; d < 3; ++d) recs[i].v[d] = h->h_bc[i].v[d];
            recs[i].rho = h->h_bc[i].rho;
            if (recs[i].kind == LBM_B200_INFLOW) feq_host(h->Q, recs[i].rho, recs[i].v, recs[i].feq);
        }
        if (h->d_bc) CU(cudaFree(h->d_bc));
        h->d_bc = nullptr;
        CU(cudaMalloc(&h->d_bc, recs.size() * sizeof(BcRec)));
        CU(cudaMemcpyAsync(h->d_bc, recs.data(), recs.size() * sizeof(BcRec), cudaMemcpyHostToDevice, h->stream));
        CU(cudaStreamSynchronize(h->stream));
    }
    if (h->geom_unchecked) {
        TRY(validate_maps(h, nullptr, nullptr));
        h->geom_unchecked = false;
    }
    // ghost-shell cells that kept the fluid handler, and periodic z: only the shell is scanned
    std::vector<int> ghost;
    bool periodic_z = false;
    const bool z_lo_shell = h->z_first == 1, z_hi_shell = h->z_first + g.zl - 1 == h->zl_global;
    auto visit = [&](int x, int y, int z) {
        const int k = h->h_kind[((size_t) z * (g.yl + 2) + y) * (g.xl + 2) + x];
        if (k == K_FLUID) ghost.push_back(cell_at(g, x, y, z));
        if (k == K_PERIODIC && (z == 0 || z == g.zl + 1)) periodic_z = true;
    };
    for (int z = 0; z < g.zl + 2; ++z) {
        const bool zshell = (z == 0 && z_lo_shell) || (z == g.zl + 1 && z_hi_shell);
        if ((z == 0 || z == g.zl + 1) && !zshell) continue;   // interface ghost plane: the neighbour's cells
        for (int y = 0; y < g.yl + 2; ++y) {
            if (zshell || y == 0 || y == g.yl + 1) {
                for (int x = 0; x < g.xl + 2; ++x) visit(x, y, z);
            } else {
                visit(0, y, z);
                visit(g.xl + 1, y, z);
            }
        }
    }
    h->ring_lo = h->ring_hi = false;
    if (periodic_z && (h->z_first != 1 || g.zl != h->zl_global)) {
        // closed by a ring of slabs instead: the first and the last slab must be each other's neighbours
        h->ring_lo = h->z_first == 1;
        h->ring_hi = h->z_first + g.zl - 1 == h->zl_global;
        periodic_z = false;
    }
    h->wrap_z = periodic_z ? 1 : 0;
    if (!ghost.empty() && g.zl != h->zl_global)
        return fail(LBM_B200_EINVAL, "%zu ghost-shell cells carry the fluid handler; on a multi-slab domain the "
                    "whole shell must be covered by boundary conditions", ghost.size());

    // dense -> padded maps through a device staging copy
    {
        DevBuf buf8, buf16;          // own the staging for the duration of this block either way
        buf8.p = h->stage_kind;
        buf16.p = h->stage_bcid;
        h->stage_kind = nullptr;
        h->stage_bcid = nullptr;
        if (!buf8.p || !buf16.p) {   // maps edited on the host (boxes, mask): upload the host copies
            if (buf8.p) { cudaFree(buf8.p); buf8.p = nullptr; }
            if (buf16.p) { cudaFree(buf16.p); buf16.p = nullptr; }
            CU(cudaMalloc(&buf8.p, n));
            CU(cudaMalloc(&buf16.p, n * sizeof(uint16_t)));
            CU(cudaMemcpyAsync(buf8.p, h->h_kind.data(), n, cudaMemcpyHostToDevice, h->stream));
            CU(cudaMemcpyAsync(buf16.p, h->h_bcid.data(), n * sizeof(uint16_t), cudaMemcpyHostToDevice, h->stream));
        }
        uint8_t* stage8 = buf8.as<uint8_t>();
        uint16_t* stage16 = buf16.as<uint16_t>();
        CU(cudaMemsetAsync(h->d_kind, K_NULL, h->map_elems(), h->stream));
        CU(cudaMemsetAsync(h->d_bcid, 0, h->map_elems() * sizeof(uint16_t), h->stream));
        scatter_map_kernel<uint8_t><<<map_grid(g, g.zl + 2), 128, 0, h->stream>>>(stage8, h->d_kind, g, 0);
        scatter_map_kernel<uint16_t><<<map_grid(g, g.zl + 2), 128, 0, h->stream>>>(stage16, h->d_bcid, g, 0);
        CU(cudaMemsetAsync(h->d_mask, 0x80, h->map_elems() * sizeof(uint32_t), h->stream));
        CU(cudaMemsetAsync(h->d_bits, 0, (h->map_elems() / 32 + 2) * sizeof(uint32_t), h->stream));
        dispatch_q(h->Q, [&](auto Qc) {
            build_mask_kernel<decltype(Qc)::value><<<map_grid(g, g.zl + 2), 128, 0, h->stream>>>(h->d_kind, h->d_mask, h->d_bits, g);
            return 0;
        });
        h->launches += 3;
        CU(cudaGetLastError());
        CU(cudaStreamSynchronize(h->stream));
    }
    if (h->d_ghost) CU(cudaFree(h->d_ghost));
    h->d_ghost = nullptr;
    h->n_ghost = (int) ghost.size();
    if (h->n_ghost) {
        CU(cudaMalloc(&h->d_ghost, ghost.size() * sizeof(int)));
        CU(cudaMemcpy(h->d_ghost, ghost.data(), ghost.size() * sizeof(int), cudaMemcpyHostToDevice));
    }
    h->geom_dirty = false;
    return 0;
}

int launch_sweep(lbm_b200* h, int z0, int nz, bool with_peers, int z_step = 1)
{
    if (nz <= 0) return 0;
    const Layout& g = h->g;
    SweepParams p{};
    p.src = h->f[h->cur];
    p.dst = h->f[1 - h->cur];
    p.mask = h->d_mask;
    p.bits = h->d_bits;
    p.kind = h->d_kind;
    p.bcid = h->d_bcid;
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
    dim3 grid((g.xl + bx - 1) / bx, (g.yl + by - 1) / by, nz);
    dispatch_q(h->Q, [&](auto Qc) {
        constexpr int Q = decltype(Qc)::value;
        if (h->exact) sweep_kernel<Q, true><<<grid, LBM_SWEEP_THREADS, 0, h->stream>>>(p);
        else sweep_kernel<Q, false><<<grid, LBM_SWEEP_THREADS, 0, h->stream>>>(p);
        return 0;
    });
    h->launches++;
    CU(cudaGetLastError());
    return 0;
}

int launch_ghost(lbm_b200* h)
{
    if (!h->n_ghost) return 0;
    double* field = h->f[1 - h->cur];
    dispatch_q(h->Q, [&](auto Qc) {
        constexpr int Q = decltype(Qc)::value;
        const int blocks = (h->n_ghost + 127) / 128;
        if (h->exact) ghost_fluid_kernel<Q, true><<<blocks, 128, 0, h->stream>>>(field, h->g.qstride, h->d_ghost, h->n_ghost, h->tau, 1.0 / h->tau);
        else ghost_fluid_kernel<Q, false><<<blocks, 128, 0, h->stream>>>(field, h->g.qstride, h->d_ghost, h->n_ghost, h->tau, 1.0 / h->tau);
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
    const long long timeout = (long long) h->clock_khz * 1000 * 20;    // ~20 s
    halo_wait_kernel<<<1, 1, 0, h->stream>>>(h->peer_flag[LBM_B200_DOWN] ? h->d_flags + LBM_B200_DOWN : nullptr,
                                             h->peer_flag[LBM_B200_UP] ? h->d_flags + LBM_B200_UP : nullptr,
                                             h->halo_epoch, timeout, h->d_halo_error,
                                             h->halo_epoch < lbm_b200::TRACE_EPOCHS ? h->d_trace : nullptr);
    h->launches++;
    CU(cudaGetLastError());
    return 0;
}
int halo_signal(lbm_b200* h)
{
    if (!has_peers(h)) return 0;
    h->halo_epoch++;
    halo_signal_kernel<<<1, 1, 0, h->stream>>>(h->peer_flag[LBM_B200_DOWN], h->peer_flag[LBM_B200_UP], h->halo_epoch, h->d_flags + 4);
    h->launches++;
    CU(cudaGetLastError());
    return 0;
}
int halo_check(lbm_b200* h)
{
    if (!has_peers(h)) return 0;
    int err = 0;
    CU(cudaMemcpyAsync(&err, h->d_halo_error, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    if (err) return fail(LBM_B200_ETIMEOUT, "a neighbour slab did not complete its sweep within the hand-shake timeout");
    return 0;
}

void finish_step(lbm_b200* h)
{
    h->cur = 1 - h->cur;      // Domain::swap, domain.hpp:169-172
    h->first = false;
    h->materialized = false;
    h->steps++;
}

int materialize(lbm_b200* h)
{
    TRY(commit_geometry(h));
    if (h->materialized) return 0;
    const Layout& g = h->g;
    // interface ghost planes count as interior only if the neighbours' full planes were pushed
    // there for this time level (lbm_b200_halo_push_all); otherwise links across a cut are skipped
    const int lo_open = h->z_first != 1 && h->full_halo_at == h->steps;
    const int hi_open = h->z_first + g.zl - 1 != h->zl_global && h->full_halo_at == h->steps;
    dispatch_q(h->Q, [&](auto Qc) {
        constexpr int Q = decltype(Qc)::value;
        if (h->exact) materialize_kernel<Q, true><<<map_grid(g, g.zl + 2), 128, 0, h->stream>>>(h->f[h->cur], h->d_kind, h->d_bcid, h->d_bc, g, lo_open, hi_open);
        else materialize_kernel<Q, false><<<map_grid(g, g.zl + 2), 128, 0, h->stream>>>(h->f[h->cur], h->d_kind, h->d_bcid, h->d_bc, g, lo_open, hi_open);
        return 0;
    });
    h->launches++;
    CU(cudaGetLastError());
    h->materialized = true;
    return 0;
}

void drop_staged_maps(lbm_b200* h)
{
    if (h->stage_kind) cudaFree(h->stage_kind);
    if (h->stage_bcid) cudaFree(h->stage_bcid);
    h->stage_kind = nullptr;
    h->stage_bcid = nullptr;
}

// Geometry edits after time steps: first let the boundary cells of the OLD geometry take the values
// the reference holds (its non-fluid pass ran after every step), then remember that the next stream
// must pull stored values -- new boundary cells still carry their former fluid populations.
int before_geometry_change(lbm_b200* h)
{
    if (h->steps > 0 && !h->geom_dirty) {
        DeviceGuard guard(h->device);
        if (!guard.ok) return fail(LBM_B200_ECUDA, "cannot select CUDA device %d", h->device);
        TRY(materialize(h));
    }
    h->first = true;
    return 0;
}

int create_common(lbm_b200_t** out, int Q, uint64_t xl, uint64_t yl, uint64_t zl_global, uint64_t z_first,
                  uint64_t zl_local, double tau, int device)
{
    if (!out) return fail(LBM_B200_EINVAL, "null output pointer");
    *out = nullptr;
    if (Q != 15 && Q != 19 && Q != 27) return fail(LBM_B200_EINVAL, "Q must be 15, 19 or 27 (got %d)", Q);
    if (xl == 0 || yl == 0 || zl_global == 0 || zl_local == 0)
        return fail(LBM_B200_EINVAL, "domain lengths must be positive");
    if (z_first < 1 || z_first + zl_local - 1 > zl_global)
        return fail(LBM_B200_EINVAL, "slab [%llu, %llu] outside 1..%llu", (unsigned long long) z_first,
                    (unsigned long long) (z_first + zl_local - 1), (unsigned long long) zl_global);
    if (!(tau > 0.0)) return fail(LBM_B200_EINVAL, "tau must be positive (got %g)", tau);
    if (yl + 2 > 65535 || zl_local + 2 > 65535) return fail(LBM_B200_EINVAL, "yl and zl are limited to 65533");
    const uint64_t P = (xl + 2 + 15) / 16 * 16;
    const uint64_t plane = P * (yl + 2);
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
    h->g.xl = (int) xl; h->g.yl = (int) yl; h->g.zl = (int) zl_local;
    h->g.P = (int) P; h->g.plane = (int) plane; h->g.qstride = (long long) qstride;
    h->zl_global = (int) zl_global;
    h->z_first = (int) z_first;
    h->tau = tau;
    {
        double vel[27 * 3];
        lbm_b200_model(Q, vel, nullptr);
        for (int q = 0; q < Q; ++q)
            h->pull_offset[q] = (long long) vel[3 * q + 2] * h->g.plane + (long long) vel[3 * q + 1] * h->g.P + (long long) vel[3 * q];
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
    CUB(cudaEventCreate(&h->ev_a));
    CUB(cudaEventCreate(&h->ev_b));
    // one allocation for both lattices: a neighbour maps it with a single IPC handle
    // (IPC handles address whole allocations, so the hand-shake counters live in its tail)
    CUB(cudaMalloc(&h->f[0], 2 * h->field_bytes() + 256));
    h->f[1] = h->f[0] + (size_t) h->g.qstride * Q;
    h->d_flags = reinterpret_cast<unsigned long long*>(h->f[0] + 2 * (size_t) h->g.qstride * Q);
    CUB(cudaMemset(h->d_flags, 0, 256));
    CUB(cudaMalloc(&h->d_mask, h->map_elems() * sizeof(uint32_t)));
    CUB(cudaMalloc(&h->d_bits, (h->map_elems() / 32 + 2) * sizeof(uint32_t)));
    CUB(cudaMalloc(&h->d_kind, h->map_elems()));
    CUB(cudaMalloc(&h->d_bcid, h->map_elems() * sizeof(uint16_t)));
    if (cudaDeviceGetAttribute(&h->clock_khz, cudaDevAttrClockRate, device) != cudaSuccess || h->clock_khz <= 0) {
        cudaGetLastError();
        h->clock_khz = 1965000;
    }
    CUB(cudaMalloc(&h->d_halo_error, sizeof(int)));
    CUB(cudaMemset(h->d_halo_error, 0, sizeof(int)));
    if (getenv("LBM_B200_HALO_TRACE")) {
        CUB(cudaMalloc(&h->d_trace, 2 * lbm_b200::TRACE_EPOCHS * sizeof(unsigned long long)));
        CUB(cudaMemset(h->d_trace, 0, 2 * lbm_b200::TRACE_EPOCHS * sizeof(unsigned long long)));
    }
#undef CUB
    h->h_kind.assign(h->ncell(), (uint8_t) LBM_B200_FLUID);   // domain.hpp:87-93
    h->h_bcid.assign(h->ncell(), 0);
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
    return create_common(h, Q, xl, yl, zl, 1, zl, tau, device);
}
int lbm_b200_create_slab(lbm_b200_t** h, int Q, uint64_t xl, uint64_t yl, uint64_t zl_global, uint64_t z_first,
                         uint64_t zl_local, double tau, int device)
{
    return create_common(h, Q, xl, yl, zl_global, z_first, zl_local, tau, device);
}

int lbm_b200_destroy(lbm_b200_t* h)
{
    if (!h) return 0;
    DeviceGuard guard(h->device);
    if (h->own_stream) cudaStreamSynchronize(h->own_stream);
    if (h->d_trace) {   // diagnostic dump: "<prefix>.<device>.z<first plane>" with entry/exit ns per epoch
        std::vector<unsigned long long> t(2 * lbm_b200::TRACE_EPOCHS);
        cudaDeviceSynchronize();
        if (cudaMemcpy(t.data(), h->d_trace, t.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost) == cudaSuccess) {
            const std::string path = std::string(getenv("LBM_B200_HALO_TRACE") ? getenv("LBM_B200_HALO_TRACE") : "halo_trace")
                    + "." + std::to_string(h->device) + ".z" + std::to_string(h->z_first);
            if (FILE* fp = fopen(path.c_str(), "w")) {
                for (unsigned long long e = 0; e < h->halo_epoch && e < (unsigned long long) lbm_b200::TRACE_EPOCHS; ++e)
                    fprintf(fp, "%llu %llu %llu\n", e, t[2 * e], t[2 * e + 1]);
                fclose(fp);
            }
        }
        cudaFree(h->d_trace);
    }
    if (h->peer_ipc_shared) h->peer_ipc_base[1] = nullptr;
    for (int s = 0; s < 2; ++s) {
        if (h->peer_ipc_base[s]) cudaIpcCloseMemHandle(h->peer_ipc_base[s]);
    }
    if (h->d_halo_error) cudaFree(h->d_halo_error);
    if (h->f[0]) cudaFree(h->f[0]);
    if (h->d_mask) cudaFree(h->d_mask);
    if (h->d_bits) cudaFree(h->d_bits);
    if (h->d_kind) cudaFree(h->d_kind);
    if (h->d_bcid) cudaFree(h->d_bcid);
    if (h->d_bc) cudaFree(h->d_bc);
    if (h->d_ghost) cudaFree(h->d_ghost);
    drop_staged_maps(h);
    if (h->d_rho) cudaFree(h->d_rho);
    if (h->d_u) cudaFree(h->d_u);
    if (h->ev_a) cudaEventDestroy(h->ev_a);
    if (h->ev_b) cudaEventDestroy(h->ev_b);
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
    return 0;
}
int lbm_b200_set_tau(lbm_b200_t* h, double tau)
{
    if (!h) return fail(LBM_B200_EINVAL, "null handle");
    if (!(tau > 0.0)) return fail(LBM_B200_EINVAL, "tau must be positive (got %g)", tau);
    h->tau = tau;
    return 0;
}
int lbm_b200_set_stream(lbm_b200_t* h, void* cuda_stream)
{
    GUARD(h);
    CU(cudaStreamSynchronize(h->stream));
    h->stream = cuda_stream ? (cudaStream_t) cuda_stream : h->own_stream;
    return 0;
}

int lbm_b200_set_geometry(lbm_b200_t* h, const uint8_t* kind, const uint16_t* bc_id, const lbm_b200_bc* table, int n_table)
{
    GUARD(h);
    if (!kind) return fail(LBM_B200_EINVAL, "kind map is null");
    if (n_table < 0 || n_table > 65535 || (n_table > 0 && !table)) return fail(LBM_B200_EINVAL, "bad boundary table");
    if (!bc_id && n_table > 1) return fail(LBM_B200_EINVAL, "bc_id map required for a table of %d handlers", n_table);
    TRY(before_geometry_change(h));
    const size_t n = h->ncell();
    // the caller's arrays go to the device straight away (fast when they are pinned) while the host
    // copies are taken and validated in parallel; commit_geometry() later only scatters them
    drop_staged_maps(h);
    CU(cudaMalloc(&h->stage_kind, n));
    CU(cudaMalloc(&h->stage_bcid, n * sizeof(uint16_t)));
    CU(cudaMemcpyAsync(h->stage_kind, kind, n, cudaMemcpyHostToDevice, h->stream));
    if (bc_id) CU(cudaMemcpyAsync(h->stage_bcid, bc_id, n * sizeof(uint16_t), cudaMemcpyHostToDevice, h->stream));
    else CU(cudaMemsetAsync(h->stage_bcid, 0, n * sizeof(uint16_t), h->stream));
    h->h_bc.assign(table, table + n_table);
    h->h_kind.resize(n);
    h->h_bcid.resize(n);
    h->geom_dirty = true;
    h->materialized = false;
    const int rc = validate_maps(h, kind, bc_id);
    CU(cudaStreamSynchronize(h->stream));     // the caller may release its arrays after this call
    if (rc != 0) {
        drop_staged_maps(h);
        h->geom_unchecked = true;             // the next commit reports the same error again
        return rc;
    }
    h->geom_unchecked = false;
    return 0;
}

int lbm_b200_set_boxes(lbm_b200_t* h, const uint64_t* boxes6, const lbm_b200_bc* table, int n)
{
    if (!h) return fail(LBM_B200_EINVAL, "null handle");
    if (n < 0 || (n > 0 && (!boxes6 || !table))) return fail(LBM_B200_EINVAL, "bad box list");
    TRY(before_geometry_change(h));
    drop_staged_maps(h);   // the host maps become the source of the next commit
    const Layout& g = h->g;
    for (int b = 0; b < n; ++b) {
        const uint64_t* e = boxes6 + 6 * b;
        // the reference asserts these (domain.hpp:180-181)
        if (!(e[1] >= e[0] && e[3] >= e[2] && e[5] >= e[4]))
            return fail(LBM_B200_EINVAL, "box %d: end before begin", b);
        if (!(e[1] < (uint64_t) g.xl + 2 && e[3] < (uint64_t) g.yl + 2 && e[5] < (uint64_t) h->zl_global + 2))
            return fail(LBM_B200_EINVAL, "box %d: extent outside the domain", b);
        const int k = table[b].kind;
        if (!((k >= K_NOSLIP && k <= K_PRESSURE) || k == K_PARALLEL || k == K_PERIODIC || k == K_NULL || k == K_FLUID))
            return fail(LBM_B200_EINVAL, "box %d: unknown kind %d", b, k);
        if (h->h_bc.size() >= 65535) return fail(LBM_B200_EINVAL, "more than 65535 boundary handlers");
        const uint16_t id = (uint16_t) h->h_bc.size();
        h->h_bc.push_back(table[b]);
        const long long zoff = h->z_first - 1;   // local z = global z - zoff
        const long long lz0 = std::max<long long>((long long) e[4] - zoff, 0);
        const long long lz1 = std::min<long long>((long long) e[5] - zoff, g.zl + 1);
        for (long long z = lz0; z <= lz1; ++z)
            for (uint64_t y = e[2]; y <= e[3]; ++y) {
                const size_t row = ((size_t) z * (g.yl + 2) + y) * (g.xl + 2);
                for (uint64_t x = e[0]; x <= e[1]; ++x) {
                    h->h_kind[row + x] = (uint8_t) k;
                    h->h_bcid[row + x] = id;
                }
            }
    }
    h->geom_dirty = true;
    h->materialized = false;
    return 0;
}

int lbm_b200_set_fluid_mask(lbm_b200_t* h, const uint8_t* mask)
{
    if (!h) return fail(LBM_B200_EINVAL, "null handle");
    if (!mask) return fail(LBM_B200_EINVAL, "mask is null");
    if (h->h_bc.size() >= 65535) return fail(LBM_B200_EINVAL, "more than 65535 boundary handlers");
    TRY(before_geometry_change(h));
    drop_staged_maps(h);
    const Layout& g = h->g;
    lbm_b200_bc solid{};
    solid.kind = LBM_B200_NOSLIP;
    solid.rho = 1.0;
    const uint16_t id = (uint16_t) h->h_bc.size();
    h->h_bc.push_back(solid);
    size_t i = 0;
    for (int z = 1; z < g.zl + 1; ++z)
        for (int y = 1; y < g.yl + 1; ++y) {
            const size_t row = ((size_t) z * (g.yl + 2) + y) * (g.xl + 2);
            for (int x = 1; x < g.xl + 1; ++x, ++i)
                if (!mask[i]) {
                    h->h_kind[row + x] = LBM_B200_NOSLIP;
                    h->h_bcid[row + x] = id;
                }
        }
    h->geom_dirty = true;
    h->materialized = false;
    return 0;
}

int lbm_b200_get_kind(lbm_b200_t* h, uint8_t* kind)
{
    if (!h || !kind) return fail(LBM_B200_EINVAL, "null argument");
    memcpy(kind, h->h_kind.data(), h->ncell());
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
// File: 64-byte header {magic "LBMB200\0", version, Q, xl, yl, zl_local, z_first, zl_global, steps, tau}
// followed by the collide field as Q dense arrays in Domain::idx order (layout LBM_B200_SOA).
namespace {
struct CheckpointHeader {
    char magic[8];
    int32_t version, Q;
    int32_t xl, yl, zl_local, z_first, zl_global, pad;
    uint64_t steps;
    double tau;
    uint64_t reserved;
};
static_assert(sizeof(CheckpointHeader) == 64, "checkpoint header layout");
}

int lbm_b200_save_checkpoint(lbm_b200_t* h, const char* path)
{
    GUARD(h);
    if (!path) return fail(LBM_B200_EINVAL, "null path");
    TRY(materialize(h));
    const Layout& g = h->g;
    FILE* fp = fopen(path, "wb");
    if (!fp) return fail(LBM_B200_EINVAL, "cannot open %s for writing", path);
    CheckpointHeader hd{};
    memcpy(hd.magic, "LBMB200", 8);
    hd.version = 1; hd.Q = h->Q;
    hd.xl = g.xl; hd.yl = g.yl; hd.zl_local = g.zl; hd.z_first = h->z_first; hd.zl_global = h->zl_global;
    hd.steps = h->steps; hd.tau = h->tau;
    int rc = fwrite(&hd, sizeof hd, 1, fp) == 1 ? 0 : fail(LBM_B200_EINVAL, "write to %s failed", path);
    const size_t n = h->ncell();
    std::vector<double> buf(n);
    const size_t rows = (size_t) (g.yl + 2) * (g.zl + 2);
    for (int q = 0; q < h->Q && rc == 0; ++q) {
        const double* d = h->f[h->cur] + (size_t) q * g.qstride + X_SHIFT;
        cudaError_t e = cudaMemcpy2DAsync(buf.data(), (g.xl + 2) * sizeof(double), d, g.P * sizeof(double),
                                          (g.xl + 2) * sizeof(double), rows, cudaMemcpyDeviceToHost, h->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
        if (e != cudaSuccess) rc = fail(LBM_B200_ECUDA, "checkpoint download failed: %s", cudaGetErrorString(e));
        else if (fwrite(buf.data(), sizeof(double), n, fp) != n) rc = fail(LBM_B200_EINVAL, "write to %s failed", path);
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
    if (fread(&hd, sizeof hd, 1, fp) != 1 || memcmp(hd.magic, "LBMB200", 8) != 0 || hd.version != 1)
        rc = fail(LBM_B200_EINVAL, "%s is not a lbm_b200 checkpoint", path);
    else if (hd.Q != h->Q || hd.xl != g.xl || hd.yl != g.yl || hd.zl_local != g.zl || hd.z_first != h->z_first || hd.zl_global != h->zl_global)
        rc = fail(LBM_B200_EINVAL, "%s holds D3Q%d %dx%dx%d (slab at %d of %d), this domain is D3Q%d %dx%dx%d (slab at %d of %d)", path,
                  hd.Q, hd.xl, hd.yl, hd.zl_local, hd.z_first, hd.zl_global, h->Q, g.xl, g.yl, g.zl, h->z_first, h->zl_global);
    const size_t n = h->ncell();
    std::vector<double> buf(rc == 0 ? n : 0);
    const size_t rows = (size_t) (g.yl + 2) * (g.zl + 2);
    for (int q = 0; q < h->Q && rc == 0; ++q) {
        if (fread(buf.data(), sizeof(double), n, fp) != n) { rc = fail(LBM_B200_EINVAL, "%s is truncated", path); break; }
        double* d = h->f[h->cur] + (size_t) q * g.qstride + X_SHIFT;
        cudaError_t e = cudaMemcpy2DAsync(d, g.P * sizeof(double), buf.data(), (g.xl + 2) * sizeof(double),
                                          (g.xl + 2) * sizeof(double), rows, cudaMemcpyHostToDevice, h->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
        if (e != cudaSuccess) rc = fail(LBM_B200_ECUDA, "checkpoint upload failed: %s", cudaGetErrorString(e));
    }
    fclose(fp);
    if (rc == 0) {
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
    if (h->edges_done) return fail(LBM_B200_ESTATE, "a split-phase step is in flight");
    TRY(commit_geometry(h));
    if ((h->ring_lo && !h->peer_f[LBM_B200_DOWN][0]) || (h->ring_hi && !h->peer_f[LBM_B200_UP][0]))
        return fail(LBM_B200_ESTATE, "periodic z on a multi-slab domain needs the first and last slab connected as a ring "
                    "(lbm_b200_connect / lbm_b200_connect_local on that side)");
    CU(cudaEventRecord(h->ev_a, h->stream));
    for (uint64_t s = 0; s < n_steps; ++s) {
        if (has_peers(h) && h->g.zl >= 3) {
            // Only the two edge planes read ghost planes and feed the neighbours, so only they take
            // part in the hand-shake; the interior sweep that follows gives every neighbour a whole
            // step of slack before its next wait.
            TRY(halo_wait(h));
            TRY(launch_sweep(h, 1, 2, true, h->g.zl - 1));
            TRY(halo_signal(h));
            TRY(launch_sweep(h, 2, h->g.zl - 2, false));
        } else {
            TRY(halo_wait(h));
            TRY(launch_sweep(h, 1, h->g.zl, true));
            TRY(halo_signal(h));
        }
        TRY(launch_ghost(h));
        finish_step(h);
    }
    CU(cudaEventRecord(h->ev_b, h->stream));
    h->timed = true;
    return 0;
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
uint64_t lbm_b200_steps_done(lbm_b200_t* h) { return h ? h->steps : 0; }

// ---- read-out -----------------------------------------------------------------
int lbm_b200_macroscopic(lbm_b200_t* h, double* rho, double* u)
{
    GUARD(h);
    TRY(materialize(h));
    const Layout& g = h->g;
    const size_t n = (size_t) g.xl * g.yl * g.zl;
    if (rho && !h->d_rho) CU(cudaMalloc(&h->d_rho, n * sizeof(double)));
    if (u && !h->d_u) CU(cudaMalloc(&h->d_u, 3 * n * sizeof(double)));
    double* d_rho = rho ? h->d_rho : nullptr;
    double* d_u = u ? h->d_u : nullptr;
    dim3 grid((g.xl + 127) / 128, g.yl, g.zl);
    dispatch_q(h->Q, [&](auto Qc) {
        macroscopic_kernel<decltype(Qc)::value><<<grid, 128, 0, h->stream>>>(h->f[h->cur], g, d_rho, d_u);
        return 0;
    });
    h->launches++;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess && rho) e = cudaMemcpyAsync(rho, d_rho, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess && u) e = cudaMemcpyAsync(u, d_u, 3 * n * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    if (e != cudaSuccess) return fail(LBM_B200_ECUDA, "macroscopic read-out failed: %s", cudaGetErrorString(e));
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
        diagnostics_kernel<decltype(Qc)::value><<<grid, 128, 0, h->stream>>>(h->f[h->cur], h->d_kind, g, d_part);
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
            if (L::cz(q) == want) { if (c == k) q_found = q; ++c; }
        return 0;
    });
    if (q_found < 0) return fail(LBM_B200_EINVAL, "halo population index %d out of range", k);
    const Layout& g = h->g;
    int z;
    if (side == LBM_B200_UP) z = recv ? g.zl + 1 : g.zl;
    else z = recv ? 0 : 1;
    *ptr = h->f[buffer] + (size_t) q_found * g.qstride + (size_t) z * g.plane;
    return 0;
}

int lbm_b200_dst_buffer(lbm_b200_t* h) { return h ? 1 - h->cur : -1; }

int lbm_b200_step_edges(lbm_b200_t* h)
{
    GUARD(h);
    if (h->edges_done) return fail(LBM_B200_ESTATE, "step_edges called twice");
    TRY(commit_geometry(h));
    if (h->g.zl > 1) TRY(launch_sweep(h, 1, 2, true, h->g.zl - 1));
    else TRY(launch_sweep(h, 1, 1, true));
    h->edges_done = true;
    return 0;
}
int lbm_b200_step_interior(lbm_b200_t* h)
{
    GUARD(h);
    if (!h->edges_done) return fail(LBM_B200_ESTATE, "step_interior before step_edges");
    TRY(launch_sweep(h, 2, h->g.zl - 2, true));
    TRY(launch_ghost(h));
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
        const int z_mine = side == LBM_B200_UP ? g.zl : 1;
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
    h->materialized = false;
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
    long long meta[4] = { h->g.qstride, h->g.plane, h->g.zl, h->Q };
    memcpy(p + 64, meta, sizeof meta);
    return 0;
}

static int connect_common(lbm_b200* h, int side, double* base, unsigned long long* nb_flags, long long qstride,
                          long long plane, long long zl, long long Q)
{
    if (Q != h->Q || plane != h->g.plane) return fail(LBM_B200_EINVAL, "neighbour slab has a different lattice or x-y shape");
    // we are the neighbour's DOWN side when it is our UP side, and vice versa
    h->peer_flag[side] = nb_flags + (side == LBM_B200_UP ? LBM_B200_DOWN : LBM_B200_UP);
    h->peer_f[side][0] = base;
    h->peer_f[side][1] = base + (size_t) qstride * Q;
    h->peer_qstride[side] = qstride;
    // my top plane feeds the upper neighbour's ghost plane 0; my bottom plane the
    // lower neighbour's ghost plane zl_nb+1
    h->peer_off[side] = side == LBM_B200_UP ? 0 : (zl + 1) * plane;
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
    return connect_common(h, side, nb->f[0], nb->d_flags, nb->g.qstride, nb->g.plane, nb->g.zl, nb->Q);
}

} // extern "C"

/*
 * include/lbm_b200.h -- C ABI of the B200-native lattice-Boltzmann sweep.
 *
 * This is the drop-in boundary for the hot path of hackerbruecke/lbm: the
 * three calls  domain->stream(); domain->swap(); domain->collide();  of
 * src/main.cpp:50-52, together with the state they act on (the two lattices
 * of include/domain.h:13-14) and the read-out of io/vtk.hpp:62-73.
 * The reference has no FFI of its own -- its operator API is the C++ template
 * surface (Collision/Domain/Cell) -- so these entry points are what the
 * C++ wrappers under include/lbm/ (same class names as the reference) bind to.
 * Plain pointers and sizes only; no CUDA, torch or C++ types.
 *
 * Conventions
 *   - every function returns 0 on success, a negative LBM_B200_E* code on
 *     failure; lbm_b200_last_error() then holds a message (thread-local).
 *     Nothing throws across this boundary.  (Reference: exceptions caught in
 *     main, src/main.cpp:68-71; the C++ wrapper re-throws.)
 *   - host arrays over cells use the reference's Domain::idx order
 *     (domain.hpp:61-64):  idx = x + (xl+2)*y + (xl+2)*(yl+2)*z, ghost shell
 *     included, local to the handle (a slab has zl_local+2 planes).
 *   - there is NO CPU fallback: without a CUDA device every compute entry
 *     point fails with LBM_B200_ECUDA.
 *   - a handle is driven by one host thread at a time (like the reference's
 *     Domain); different handles may be used from different threads.  Calls
 *     that enqueue work (step) return before the GPU has finished; calls that
 *     return data synchronise.
 */
#ifndef LBM_B200_H
#define LBM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LBM_B200_ABI_VERSION 2

/* error codes */
enum {
    LBM_B200_OK = 0,
    LBM_B200_EINVAL = -1,   /* bad argument                                   */
    LBM_B200_ECUDA = -2,    /* CUDA runtime error / no device                  */
    LBM_B200_ENOMEM = -3,   /* allocation failed                               */
    LBM_B200_ESTATE = -4,   /* call not valid in the current state             */
    LBM_B200_ETIMEOUT = -5  /* neighbour slab did not signal in time           */
};

/* Cell handler kinds == the reference's collision classes. */
enum {
    LBM_B200_FLUID = 0,      /* BGKCollision            collision.h:64-71      */
    LBM_B200_NOSLIP = 1,     /* NoSlipBoundary          boundary.hpp:15-31     */
    LBM_B200_MOVINGWALL = 2, /* MovingWallBoundary      boundary.hpp:44-68     */
    LBM_B200_FREESLIP = 3,   /* FreeSlipBoundary        boundary.hpp:80-115    */
    LBM_B200_OUTFLOW = 4,    /* OutflowBoundary         boundary.hpp:129-150   */
    LBM_B200_INFLOW = 5,     /* InflowBoundary          boundary.hpp:165-181   */
    LBM_B200_PRESSURE = 6,   /* PressureBoundary        boundary.hpp:195-214   */
    LBM_B200_NULL = 7,       /* NullCollision           collision.h:74-86      */
    LBM_B200_PARALLEL = 8,   /* parallel::ParallelBoundary (no-op) parallel.h:11-23 */
    LBM_B200_PERIODIC = 9    /* extension (not in the reference): a ghost cell  */
                             /* that mirrors its periodically wrapped interior  */
                             /* image; equals the ghost-copy recipe of SURVEY 8c */
};

/* One boundary handler object (what BoundaryKeeper::get_collision creates,
 * boundary.h:86-91): kind + constructor arguments. */
typedef struct {
    int32_t kind;
    int32_t _pad;
    double v[3];   /* wall_velocity (boundary.h:23) / inflow_velocity (boundary.h:54) */
    double rho;    /* reference_density (boundary.h:44,53) / input_density (boundary.h:65) */
} lbm_b200_bc;

/* arithmetic modes */
enum {
    LBM_B200_FAST = 0,  /* reciprocals + FMA contraction; differs from the       */
                        /* reference by rounding only (gate: 1e-12 relative)     */
    LBM_B200_EXACT = 1  /* the reference's expression association, no FMA        */
                        /* contraction, correctly rounded quotients (computed    */
                        /* from reciprocals, see lbm_b200_selftest_division):    */
                        /* bit-identical to the CPU build                        */
};

/* population layouts for upload/download */
enum {
    LBM_B200_AOS = 0,   /* f[idx*Q + q]  -- the reference's Cell array          */
    LBM_B200_SOA = 1    /* f[q*ncell + idx]                                     */
};
/* which of the reference's two lattices (domain.h:13-14) */
enum {
    LBM_B200_COLLIDE_FIELD = 0,  /* what Domain::cell() addresses              */
    LBM_B200_STREAM_FIELD = 1
};

typedef struct lbm_b200 lbm_b200_t;

const char* lbm_b200_last_error(void);
int lbm_b200_abi_version(void);
/* number of visible CUDA devices (0 without a GPU; never fails) */
int lbm_b200_device_count(void);

/* --- lattice descriptors (model.h:13-134), host side, no GPU needed -------- */
/* velocities: Q*3 doubles (the reference stores them as double), weights: Q  */
int lbm_b200_model(int Q, double* velocities, double* weights);
int lbm_b200_model_inv(int Q, int q);
int lbm_b200_model_velocity_index(int Q, int u, int v, int w);

/* --- life cycle ------------------------------------------------------------ */
/* Domain<M>(xl,yl,zl,collision) with BGKCollision<M>(tau)  (domain.hpp:87-98,
 * collision.hpp:55-58).  Both lattices start at the weights (cell.hpp:9-15),
 * every cell -- ghost shell included -- has the fluid handler.
 * device < 0 selects the current CUDA device. */
int lbm_b200_create(lbm_b200_t** h, int Q, uint64_t xl, uint64_t yl, uint64_t zl,
                    double tau, int device);
/* One z-slab of a global domain (replaces the intent of parallel.h:11-23 and
 * Domain::create_subdomain, domain.hpp:197-248): this handle owns the global
 * interior planes z_first .. z_first+zl_local-1 (1-based) of zl_global, plus one
 * ghost plane on each side. */
int lbm_b200_create_slab(lbm_b200_t** h, int Q, uint64_t xl, uint64_t yl, uint64_t zl_global,
                         uint64_t z_first, uint64_t zl_local, double tau, int device);
/* The same for either split axis: (xl, yl, zl) is the GLOBAL domain, the handle owns the planes first ..
 * first+local-1 (1-based) of `axis`: LBM_B200_AXIS_Z = x-y planes (z-slabs, as above), LBM_B200_AXIS_Y = x-z planes
 * (y-slabs: for domains whose z extent is shorter than the number of GPUs, or flat ones such as the reference's
 * shearflow scenario, 8 x 8 x 20).  A y-slab stores y as the slowest index, so its planes are contiguous and every
 * halo function below works unchanged; its host arrays use the local lengths (xl, local, zl) in Domain::idx order. */
enum { LBM_B200_AXIS_Y = 1, LBM_B200_AXIS_Z = 2 };
int lbm_b200_create_slab_axis(lbm_b200_t** h, int Q, uint64_t xl, uint64_t yl, uint64_t zl, int axis,
                              uint64_t first, uint64_t local, double tau, int device);
int lbm_b200_destroy(lbm_b200_t* h);

int lbm_b200_set_arithmetic(lbm_b200_t* h, int mode);
int lbm_b200_set_tau(lbm_b200_t* h, double tau);
/* run on a caller-owned CUDA stream (cudaStream_t passed as void*); NULL
 * returns to the handle's own stream */
int lbm_b200_set_stream(lbm_b200_t* h, void* cuda_stream);

/* --- geometry (Domain::setBoundaryCondition, domain.hpp:175-194;
 *     Domain::set_nonfluid_cells_nullcollide, domain.hpp:101-113; the mask loop of
 *     io/vtk.hpp:137-150) -------------------------------------------------------
 * The handler maps live on the DEVICE (one kind byte + one handler id per cell of
 * each lattice); boxes and masks are painted there by small kernels, so applying a
 * scenario to a 512^3 lattice moves a few hundred bytes over PCIe, not the maps.
 *
 * Handlers belong to lattices in the reference (cell.h:15).  setBoundaryCondition
 * writes both lattices; everything that goes through Domain::cell() -- the VTK mask
 * reader, set_nonfluid_cells_nullcollide, Cell::set_collision_handler -- writes the
 * collide field only, after which such a cell is solid on every other step.  The
 * `literal` arguments below select that behaviour (bit-identical to the reference);
 * literal = 0 writes both lattices.  Literal edits are limited to whole domains
 * (not slabs). */
/* the handler table (what BoundaryKeeper owns, boundary.h:75-91): ids stored in the
 * maps index it.  Replaces the table; ids already used by cells keep their meaning,
 * so a new table must extend the old one. */
int lbm_b200_set_handlers(lbm_b200_t* h, const lbm_b200_bc* table, int n_table);
/* setBoundaryCondition for n boxes applied in order, last writer wins
 * (io/scenario.h:91-128 -> domain.hpp:185-193): box i = 6 inclusive GLOBAL indices
 * x0,xE,y0,yE,z0,zE, painted with handler ids[i] of the table (kind = table kind). */
int lbm_b200_paint_boxes(lbm_b200_t* h, const uint64_t* boxes6, const uint16_t* ids, int n);
/* convenience: appends table[0..n) to the handler table and paints box i with it */
int lbm_b200_set_boxes(lbm_b200_t* h, const uint64_t* boxes6, const lbm_b200_bc* table, int n);
/* dense maps for every local cell (Domain::idx order): kind = one LBM_B200_* per cell,
 * bc_id = index into table for cells whose kind takes parameters (may be NULL if
 * n_table <= 1; then id 0 is used).  Replaces table and maps of both lattices; the
 * maps are checked on the device and rejected as a whole if inconsistent. */
int lbm_b200_set_geometry(lbm_b200_t* h, const uint8_t* kind, const uint16_t* bc_id,
                          const lbm_b200_bc* table, int n_table);
/* the same for the local x-y planes [z_begin, z_begin+z_count) only, against the current
 * handler table: what Domain::cell(...).set_collision_handler() edits turn into */
int lbm_b200_set_geometry_planes(lbm_b200_t* h, const uint8_t* kind, const uint16_t* bc_id,
                                 uint64_t z_begin, uint64_t z_count, int literal);
/* handler kinds / ids of the collide field as Domain::cell() reports them, planes
 * [z_begin, z_begin+z_count) (either output may be NULL) */
int lbm_b200_get_geometry_planes(lbm_b200_t* h, uint8_t* kind, uint16_t* bc_id, uint64_t z_begin, uint64_t z_count);
int lbm_b200_get_kind(lbm_b200_t* h, uint8_t* kind);   /* all local planes */
/* interior fluid mask like the POINT_DATA of a legacy-VTK file (io/vtk.hpp:137-150):
 * 0 => NoSlipBoundary.  set_fluid_mask: xl*yl*zl_local bytes for this handle's own
 * interior planes, both lattices (DESIGN.md "deviations").  _literal: the same on the
 * collide field only, exactly as io/vtk.hpp:145-146.  _global: the mask of the WHOLE
 * domain (xl*yl*zl_global bytes); a slab takes its own planes and the replicas of its
 * neighbours' edge planes from it. */
int lbm_b200_set_fluid_mask(lbm_b200_t* h, const uint8_t* mask);
int lbm_b200_set_fluid_mask_literal(lbm_b200_t* h, const uint8_t* mask);
int lbm_b200_set_fluid_mask_global(lbm_b200_t* h, const uint8_t* mask, int literal);
/* the same with a handler of the table instead of an implicit NoSlipBoundary (whole-domain mask) */
int lbm_b200_paint_mask(lbm_b200_t* h, const uint8_t* mask, uint16_t id, int literal);
/* Domain::set_nonfluid_cells_nullcollide on the device; *n_tagged = cells newly tagged.
 * The reference tags the collide field only, so Domain::cell() reports NullCollision on
 * even and the former handler on odd step counts after the call; get_geometry_planes
 * reproduces that.  No population ever depends on the tag. */
int lbm_b200_tag_null_cells(lbm_b200_t* h, int literal, uint64_t* n_tagged);

/* --- state ------------------------------------------------------------------ */
int lbm_b200_upload_populations(lbm_b200_t* h, const double* f, int layout, int field);
int lbm_b200_download_populations(lbm_b200_t* h, double* f, int layout, int field);
/* the same for the local x-y planes [z_begin, z_begin+z_count) only, AoS layout: what Domain::cell() uses
 * when a few cells of a large lattice are read or written */
int lbm_b200_upload_planes(lbm_b200_t* h, const double* f, int field, uint64_t z_begin, uint64_t z_count);
int lbm_b200_download_planes(lbm_b200_t* h, double* f, int field, uint64_t z_begin, uint64_t z_count);
/* f = feq(rho,u) per local cell from host arrays rho[ncell], u[ncell*3] (idx
 * order), evaluated on the device with compute_feq's association
 * (collision.hpp:34-51 via Cell::equilibrium, cell.hpp:55-59) */
int lbm_b200_init_equilibrium(lbm_b200_t* h, const double* rho, const double* u);

/* checkpoint / restart of one slab (not in the reference): both lattices incl. boundary
 * cells as the reference holds them, the step counter, the arithmetic mode and a checksum of the handler
 * maps.  Geometry itself is not stored: re-apply it first; loading under another geometry, lattice or
 * arithmetic mode is refused. */
int lbm_b200_save_checkpoint(lbm_b200_t* h, const char* path);
int lbm_b200_load_checkpoint(lbm_b200_t* h, const char* path);

/* --- the hot path: n x { stream(); swap(); collide(); }  (src/main.cpp:50-52) */
int lbm_b200_step(lbm_b200_t* h, uint64_t n_steps);
int lbm_b200_sync(lbm_b200_t* h);
/* GPU time of the last lbm_b200_step call (CUDA events on the handle's stream) */
int lbm_b200_elapsed_ms(lbm_b200_t* h, double* ms);
/* kernels launched by this handle so far */
int lbm_b200_launch_count(lbm_b200_t* h, uint64_t* n);
/* how many of them were launches of the TMA-fed sweep (sweep_tma_kernel) */
int lbm_b200_tma_launch_count(lbm_b200_t* h, uint64_t* n);
uint64_t lbm_b200_steps_done(lbm_b200_t* h);

/* n time steps of a whole stack of connected slabs (handles[0..n_handles), any devices) from one host
 * thread: the slabs' steps are enqueued interleaved in short runs so that no device queue ever waits
 * for work that has not been submitted yet.  What Domain::step() calls for a multi-GPU Domain. */
int lbm_b200_step_group(lbm_b200_t* const* handles, int n_handles, uint64_t n_steps);
/* replay runs of steps from CUDA graphs (1 = always, 0 = never, -1 = automatic: small lattices, where
 * the launch overhead matters) */
int lbm_b200_set_graphs(lbm_b200_t* h, int mode);
/* which kernel sweeps whole planes (results are bit-identical either way; 1 = always where possible, 0 = never,
 * -1 = automatic).  tma: the TMA-fed persistent kernel (one block per SM, shared-memory rings filled by
 * cp.async.bulk.tensor) instead of one thread per cell pulling into registers -- an opt-in engine, never chosen
 * automatically (it is slower, profiles/variants_r07_tma.txt).  checked: look at the 1-bit "not a bulk cell" map
 * before pulling instead of together with the pulls -- automatic when more than 10 % of the interior cells are
 * solid. */
int lbm_b200_set_sweep_engine(lbm_b200_t* h, int tma, int checked);

/* --- read-out (io/vtk.hpp:62-73): interior cells, z,y,x order --------------- */
/* rho: xl*yl*zl_local doubles, u: 3x that (may each be NULL) */
int lbm_b200_macroscopic(lbm_b200_t* h, double* rho, double* u);
/* split form: _begin reduces density / velocity of the CURRENT time level on the device and starts the
 * copies to the host on a second stream; steps issued afterwards overlap with the transfer; _end waits
 * until rho / u are complete.  (One output interval of src/main.cpp:48-57 without stalling the time
 * loop.)  rho / u should be pinned memory (lbm_b200_host_alloc) for the copy to be asynchronous. */
int lbm_b200_macroscopic_begin(lbm_b200_t* h, double* rho, double* u);
int lbm_b200_macroscopic_end(lbm_b200_t* h);
/* page-locked host memory placed on the NUMA node next to `device` (device < 0: the current one) */
int lbm_b200_host_alloc(void** ptr, size_t bytes, int device);
int lbm_b200_host_free(void* ptr);
/* Self-test of the bit-identical mode's division (kernels.cuh: div_rcp): the reference divides by C_S*C_S, 2*C_S^4,
 * 2*C_S^2 (collision.hpp:47-48), tau (:68) and rho (:27-29); the device computes these quotients from a correctly
 * rounded reciprocal with two FMA corrections.  Compares n operands per divisor with IEEE division on the device;
 * *mismatches must come back 0. */
int lbm_b200_selftest_division(uint64_t n, uint64_t seed, double tau, uint64_t* mismatches);
/* pins the calling host thread to the CPUs next to `device` (no-op where the topology is unknown) */
int lbm_b200_bind_host_thread(int device);
/* reductions for physics checks: sum of density, sum of |u|^2, max |u| over
 * interior FLUID cells of this slab */
int lbm_b200_diagnostics(lbm_b200_t* h, double* mass, double* kinetic, double* umax);

/* --- multi-GPU z-slabs -------------------------------------------------------
 * After a step, slab r's top interior plane populations with c_z=+1 must appear
 * in slab r+1's bottom ghost plane and vice versa.  Two transports:
 *  (1) split-phase with an external transport (NCCL through torch.distributed):
 *        step_edges -> [send/recv the halo planes] -> step_interior -> step_finish
 *  (2) direct peer stores from the sweep kernel (CUDA P2P in one process, CUDA
 *      IPC across processes): connect once, then lbm_b200_step as usual;
 *      in one process lbm_b200_step_group drives the whole stack from one host thread.
 */
enum { LBM_B200_DOWN = 0, LBM_B200_UP = 1 };
/* number of populations crossing an interface per direction (5/5/9) and the
 * byte size of one x-y plane of one population (contiguous in device memory) */
int lbm_b200_halo_layout(lbm_b200_t* h, int* n_q, size_t* plane_bytes);
/* device pointer of the k-th plane to send to / receive from the neighbour on
 * `side`, inside physical buffer 0 or 1.  Send planes are the slab's edge
 * interior planes (c_z=+1 populations for UP, c_z=-1 for DOWN); receive planes
 * are the ghost planes (c_z=-1 populations arrive from UP, c_z=+1 from DOWN). */
int lbm_b200_halo_plane(lbm_b200_t* h, int buffer, int side, int k, int recv, void** device_ptr);
/* the same for ANY population q of the current collide field: the slab's edge plane on `side` (recv = 0) or the
 * ghost plane behind it (recv = 1), plane_bytes each -- what an external transport moves for the read-back
 * across slab cuts (every slab sends its Q edge planes to the neighbours, then calls lbm_b200_halo_pushed) */
int lbm_b200_edge_plane(lbm_b200_t* h, int side, int q, int recv, void** device_ptr);
/* physical buffer the step in flight writes (the one to exchange) */
int lbm_b200_dst_buffer(lbm_b200_t* h);
int lbm_b200_step_edges(lbm_b200_t* h);     /* sweep the two edge planes         */
int lbm_b200_step_interior(lbm_b200_t* h);  /* sweep the remaining planes        */
int lbm_b200_step_finish(lbm_b200_t* h);    /* swap (after the exchange landed)  */
/* CUDA-IPC export of this slab's memory (64-byte handles) and connection to a
 * neighbour's export; side = LBM_B200_UP / LBM_B200_DOWN */
#define LBM_B200_EXPORT_BYTES 256
int lbm_b200_export(lbm_b200_t* h, void* blob /*LBM_B200_EXPORT_BYTES*/);
int lbm_b200_connect(lbm_b200_t* h, int side, const void* blob);
/* Read-back across slab cuts: every slab pushes ALL populations of its edge planes into the
 * neighbours' ghost planes (halo_push_all), the caller synchronises all slabs, then tells every slab
 * (halo_pushed); the following download / macroscopic then materialises boundary cells next to a
 * cut exactly like the reference's non-fluid pass (domain.hpp:157-165).  halo_push_all needs connected peers;
 * with an external transport move the planes of lbm_b200_edge_plane instead. */
int lbm_b200_halo_push_all(lbm_b200_t* h);
int lbm_b200_halo_pushed(lbm_b200_t* h);
/* drop the peer mappings again: call on every slab (after a sync / barrier) before any is destroyed */
int lbm_b200_disconnect(lbm_b200_t* h);
/* same-process variant */
int lbm_b200_connect_local(lbm_b200_t* h, int side, lbm_b200_t* neighbour);

#ifdef __cplusplus
}
#endif
#endif /* LBM_B200_H */

/*
 * oracle/lbm_oracle.c -- TEST INFRASTRUCTURE ONLY.  Parity status: PINNED.
 *
 * Plain-C restatement of the hot path of hackerbruecke/lbm (the reference,
 * /root/reference): lattice descriptors, moments/equilibrium/BGK, the six
 * boundary handlers and the Domain stream / swap / collide loops.  It follows
 * the reference's TWO-PASS ordering literally (stream -> swap -> collide fluid
 * -> collide non-fluid), i.e. it is deliberately NOT the link-wise fused
 * formulation the CUDA product uses, so that it is an independent check.
 * Every function cites the reference file:line it restates.
 *
 * Pinned against (tests/test_oracle.py):
 *   - the reference's own headers compiled here (oracle/_ref/libref_lbm.so,
 *     oracle/ref_driver.cpp): bitwise equality of every population on cavity,
 *     channel, step, shear-flow, masked and periodic cases for Q = 15/19/27;
 *   - the committed fixtures tests/golden/ (.npz files) produced from that build by
 *     tests/golden/make_golden.py;
 *   - the survey's known-answer values.
 * The reference ships no tests or golden vectors of its own.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this; the product (lbm_b200/) never does.
 *
 * Build: gcc -std=c11 -O2 -fopenmp -ffp-contract=off -fPIC -shared (Makefile).
 * -ffp-contract=off matters: the reference's parity build has no FMA contraction
 * (semantics note 7).
 */
#include "oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <omp.h>

/* lbmdefinitions.h:47  `static constexpr double C_S = 0.57735026919l;` */
static const double C_S = 0.57735026919;

#define QMAX 27

typedef struct {
    int Q;
    double c[QMAX][3]; /* model.h: velocities are stored as double */
    double w[QMAX];
} model_t;

/*
 * model.h:13-46 (D3Q15), :53-87 (D3Q19), :96-134 (D3Q27).
 * All three tables are the 27 vectors of {-1,0,1}^3 enumerated z-slowest,
 * x-fastest, restricted to |c|^2 in {0,1,3} / {0,1,2} / {0,1,2,3}; the weight
 * depends on |c|^2 only.  Generated here instead of being typed in; the result is
 * compared entry by entry with the reference tables in tests/test_oracle.py.
 */
static int make_model(int Q, model_t* m)
{
    double w_by_norm[4];
    int allowed[4];
    if (Q == 15) {
        w_by_norm[0] = 16.0 / 72; w_by_norm[1] = 8.0 / 72; w_by_norm[2] = 0.0; w_by_norm[3] = 1.0 / 72;
        allowed[0] = 1; allowed[1] = 1; allowed[2] = 0; allowed[3] = 1;
    } else if (Q == 19) {
        w_by_norm[0] = 12.0 / 36; w_by_norm[1] = 2.0 / 36; w_by_norm[2] = 1.0 / 36; w_by_norm[3] = 0.0;
        allowed[0] = 1; allowed[1] = 1; allowed[2] = 1; allowed[3] = 0;
    } else if (Q == 27) {
        w_by_norm[0] = 64.0 / 216; w_by_norm[1] = 16.0 / 216; w_by_norm[2] = 4.0 / 216; w_by_norm[3] = 1.0 / 216;
        allowed[0] = 1; allowed[1] = 1; allowed[2] = 1; allowed[3] = 1;
    } else {
        return -1;
    }
    int q = 0;
    for (int cz = -1; cz <= 1; ++cz)
        for (int cy = -1; cy <= 1; ++cy)
            for (int cx = -1; cx <= 1; ++cx) {
                const int n = cx * cx + cy * cy + cz * cz;
                if (!allowed[n]) continue;
                m->c[q][0] = cx; m->c[q][1] = cy; m->c[q][2] = cz;
                m->w[q] = w_by_norm[n];
                ++q;
            }
    m->Q = Q;
    return q == Q ? 0 : -1;
}

/* model.h:38-41, :78-81, :124-127  inv(q) = Q-1-q */
static inline int inv_q(const model_t* m, int q) { return m->Q - 1 - q; }

/* model.h:43-46, :83-86, :130-133  velocity_index(u,v,w), integer arithmetic */
static int velocity_index(int Q, int u, int v, int w)
{
    if (Q == 15) return 5 * w + 2 * v + u + 7 - (w + 2) % 2 * (u + v) / 2;
    if (Q == 19) return w == 0 ? (6 + (v + 1) * 3 + u) : ((w + 1) * 7 + (v + 1) * 2 + u);
    return 9 * w + 3 * v + u + 13;
}

/* collision.hpp:7-14  compute_density: sequential sum from 0, q ascending */
static double density_of(const model_t* m, const double* f)
{
    double density = 0;
    for (int q = 0; q < m->Q; ++q) density += f[q];
    return density;
}

/* collision.hpp:17-31  compute_velocity: d outer, q inner, then three divisions */
static void velocity_of(const model_t* m, const double* f, double density, double* u)
{
    u[0] = 0.0; u[1] = 0.0; u[2] = 0.0;
    for (int d = 0; d < 3; ++d)
        for (int q = 0; q < m->Q; ++q) u[d] += f[q] * m->c[q][d];
    u[0] /= density;
    u[1] /= density;
    u[2] /= density;
}

/* collision.hpp:34-51  compute_feq; association as the C++ expression parses:
 * (w*rho) * (((1 + cu/(C_S*C_S)) + (cu*cu)/(2*C_S*C_S*C_S*C_S)) - uu/(2*C_S*C_S)) */
static void feq_of(const model_t* m, double density, const double* u, double* feq)
{
    for (int q = 0; q < m->Q; ++q) {
        double c_dot_u = 0.0;
        double u_dot_u = 0.0;
        for (int d = 0; d < 3; ++d) {
            c_dot_u += m->c[q][d] * u[d];
            u_dot_u += u[d] * u[d];
        }
        feq[q] = m->w[q] * density
                * (1 + c_dot_u / (C_S * C_S) + (c_dot_u) * (c_dot_u) / (2 * C_S * C_S * C_S * C_S)
                        - u_dot_u / (2 * C_S * C_S));
    }
}

/* collision.hpp:61-70  BGKCollision::collide: f -= (f - feq)/tau, in place */
static void bgk_collide(const model_t* m, double tau, double* f)
{
    double u[3], feq[QMAX];
    const double density = density_of(m, f);
    velocity_of(m, f, density, u);
    feq_of(m, density, u, feq);
    for (int q = 0; q < m->Q; ++q) f[q] -= (f[q] - feq[q]) / tau;
}

/* ------------------------------------------------------------------------- */
/* Domain state.  Two fields like domain.h:13-14; each Cell (cell.h:14-15) is  */
/* Q doubles plus a handler, kept here as an int per field:                    */
/*   H_FLUID, H_NULL, or the index of a handler record                          */
enum { H_FLUID = -1, H_NULL = -2 };

typedef struct {
    int kind;
    double v[3];
    double rho;
} handler_t;

typedef struct {
    model_t m;
    int xl, yl, zl;
    size_t n;            /* (xl+2)(yl+2)(zl+2) */
    double tau;
    double* f[2];        /* f[field][idx*Q+q] */
    int* h[2];           /* handler of every cell, per field */
    int collide;         /* which of the two is `collide_field` */
    handler_t* handlers;
} domain_t;

/* domain.hpp:61-64 */
static inline size_t idx(const domain_t* d, int x, int y, int z)
{
    return (size_t) x + (size_t) (d->xl + 2) * y + (size_t) (d->xl + 2) * (d->yl + 2) * z;
}
/* domain.hpp:67-70: strictly interior */
static inline int in_bounds(const domain_t* d, int x, int y, int z)
{
    return x > 0 && x < d->xl + 1 && y > 0 && y < d->yl + 1 && z > 0 && z < d->zl + 1;
}
/* Domain::cell() always addresses the collide field (domain.hpp:74-83) */
static inline double* cell(const domain_t* d, int x, int y, int z)
{
    return d->f[d->collide] + idx(d, x, y, z) * d->m.Q;
}
static inline int handler(const domain_t* d, int x, int y, int z)
{
    return d->h[d->collide][idx(d, x, y, z)];
}
static inline int is_fluid(const domain_t* d, int x, int y, int z)
{
    return handler(d, x, y, z) == H_FLUID;
}

/* domain.hpp:116-141  Domain::stream: interior fluid cells pull q from x - c_q */
static void stream(domain_t* d)
{
    const int Q = d->m.Q;
    double* dst = d->f[1 - d->collide];
    #pragma omp parallel for collapse(2)
    for (int z = 1; z < d->zl + 1; ++z)
        for (int y = 1; y < d->yl + 1; ++y)
            for (int x = 1; x < d->xl + 1; ++x) {
                if (!is_fluid(d, x, y, z)) continue;
                double* out = dst + idx(d, x, y, z) * Q;
                for (int q = 0; q < Q; ++q) {
                    const int sx = (int) (x - d->m.c[q][0]);
                    const int sy = (int) (y - d->m.c[q][1]);
                    const int sz = (int) (z - d->m.c[q][2]);
                    out[q] = cell(d, sx, sy, sz)[q];
                }
            }
}

/* boundary.hpp:15-31 / :44-68 / :80-115 / :129-150 / :165-181 / :195-214.
 * All six share the link test `in_bounds(x+c) && cell(x+c).is_fluid()`. */
static void boundary_collide(const domain_t* d, const handler_t* hd, int x, int y, int z)
{
    const model_t* m = &d->m;
    const int Q = m->Q;
    double* self = cell(d, x, y, z);
    for (int q = 0; q < Q; ++q) {
        const int dx = (int) m->c[q][0], dy = (int) m->c[q][1], dz = (int) m->c[q][2];
        if (!(in_bounds(d, x + dx, y + dy, z + dz) && is_fluid(d, x + dx, y + dy, z + dz))) continue;
        const double* nb = cell(d, x + dx, y + dy, z + dz);
        switch (hd->kind) {
        case ORC_NOSLIP:                                   /* boundary.hpp:28 */
            self[q] = nb[inv_q(m, q)];
            break;
        case ORC_MOVINGWALL: {                             /* boundary.hpp:57-65 */
            const double density = density_of(m, nb);
            double c_dot_u = 0;
            for (int k = 0; k < 3; ++k) c_dot_u += m->c[q][k] * hd->v[k];
            const double finv = nb[inv_q(m, q)];
            self[q] = finv + 2.0 * m->w[q] * density * c_dot_u / (C_S * C_S);
            break;
        }
        case ORC_FREESLIP:                                 /* boundary.hpp:98-112 */
            if (is_fluid(d, x + dx, y, z))
                self[q] = cell(d, x + dx, y, z)[velocity_index(Q, -dx, dy, dz)];
            else if (is_fluid(d, x, y + dy, z))
                self[q] = cell(d, x, y + dy, z)[velocity_index(Q, dx, -dy, dz)];
            else if (is_fluid(d, x, y, z + dz))
                self[q] = cell(d, x, y, z + dz)[velocity_index(Q, dx, dy, -dz)];
            else if (is_fluid(d, x, y + dy, z + dz))
                self[q] = cell(d, x, y + dy, z + dz)[velocity_index(Q, dx, -dy, -dz)];
            else if (is_fluid(d, x + dx, y, z + dz))
                self[q] = cell(d, x + dx, y, z + dz)[velocity_index(Q, -dx, dy, -dz)];
            else if (is_fluid(d, x + dx, y + dy, z))
                self[q] = cell(d, x + dx, y + dy, z)[velocity_index(Q, -dx, -dy, dz)];
            break;
        case ORC_OUTFLOW:                                  /* boundary.hpp:143-147 */
        case ORC_PRESSURE: {                               /* boundary.hpp:208-211 */
            double u[3], feq[QMAX];
            velocity_of(m, nb, hd->rho, u);               /* momentum / REFERENCE density */
            feq_of(m, hd->rho, u, feq);
            self[q] = feq[q] + feq[inv_q(m, q)] - nb[inv_q(m, q)];
            break;
        }
        case ORC_INFLOW: {                                 /* boundary.hpp:178 */
            double feq[QMAX];
            feq_of(m, hd->rho, hd->v, feq);
            self[q] = feq[q];
            break;
        }
        default:                                           /* parallel.h:20-22: no-op */
            break;
        }
    }
}

/* domain.hpp:144-166  Domain::collide: pass 1 fluid cells (ghost shell included),
 * pass 2 every non-fluid cell, which therefore sees post-collision fluid values */
static void collide(domain_t* d)
{
    #pragma omp parallel for collapse(2)
    for (int z = 0; z < d->zl + 2; ++z)
        for (int y = 0; y < d->yl + 2; ++y)
            for (int x = 0; x < d->xl + 2; ++x)
                if (is_fluid(d, x, y, z)) bgk_collide(&d->m, d->tau, cell(d, x, y, z));
    #pragma omp parallel for collapse(2)
    for (int z = 0; z < d->zl + 2; ++z)
        for (int y = 0; y < d->yl + 2; ++y)
            for (int x = 0; x < d->xl + 2; ++x) {
                const int h = handler(d, x, y, z);
                if (h >= 0) boundary_collide(d, &d->handlers[h], x, y, z);
                /* H_NULL: collision.h:81-85 does nothing */
            }
}

/* domain.hpp:169-172 */
static void swap_fields(domain_t* d) { d->collide = 1 - d->collide; }

/* domain.hpp:175-194  setBoundaryCondition: both fields, inclusive box */
static void set_boundary(domain_t* d, int h, const orc_box* b)
{
    for (uint64_t z = b->z0; z <= b->zE; ++z)
        for (uint64_t y = b->y0; y <= b->yE; ++y)
            for (uint64_t x = b->x0; x <= b->xE; ++x) {
                const size_t i = idx(d, (int) x, (int) y, (int) z);
                d->h[0][i] = h;
                d->h[1][i] = h;
            }
}

/* domain.hpp:101-113 + cell.hpp:75-92: interior non-fluid cells without an
 * in-bounds fluid neighbour (q loop includes the rest vector) get NullCollision,
 * on the collide field only (set through Domain::cell()). */
static void set_nonfluid_cells_nullcollide(domain_t* d)
{
    const model_t* m = &d->m;
    int* mark = (int*) calloc(d->n, sizeof(int));
    #pragma omp parallel for collapse(2)
    for (int z = 1; z < d->zl + 1; ++z)
        for (int y = 1; y < d->yl + 1; ++y)
            for (int x = 1; x < d->xl + 1; ++x) {
                int vicinity = 0;
                for (int q = 0; q < m->Q && !vicinity; ++q) {
                    const int nx = x + (int) m->c[q][0], ny = y + (int) m->c[q][1], nz = z + (int) m->c[q][2];
                    if (in_bounds(d, nx, ny, nz) && is_fluid(d, nx, ny, nz)) vicinity = 1;
                }
                /* has_fluid_vicinity is asked of EVERY interior cell
                 * (domain.hpp:108-109 has no is_fluid() guard); a fluid cell finds
                 * itself through the rest vector and so always keeps its handler */
                if (!vicinity) mark[idx(d, x, y, z)] = 1;
            }
    for (size_t i = 0; i < d->n; ++i)
        if (mark[i]) d->h[d->collide][i] = H_NULL;
    free(mark);
}

static int kind_of(const domain_t* d, int h)
{
    if (h == H_FLUID) return ORC_FLUID;
    if (h == H_NULL) return ORC_NULL;
    return d->handlers[h].kind;
}

static inline int wrap(int v, int l) { return v < 1 ? v + l : (v > l ? v - l : v); }

int oracle_run(const orc_case* c, orc_result* r)
{
    domain_t d;
    memset(&d, 0, sizeof d);
    if (make_model(c->Q, &d.m) != 0) return -1;
    const int Q = c->Q;
    d.xl = (int) c->xl; d.yl = (int) c->yl; d.zl = (int) c->zl;
    d.n = (size_t) (d.xl + 2) * (d.yl + 2) * (d.zl + 2);
    d.tau = c->tau;
    omp_set_num_threads(c->threads > 0 ? c->threads : 1);
    for (int k = 0; k < 2; ++k) {
        d.f[k] = (double*) malloc(d.n * Q * sizeof(double));
        d.h[k] = (int*) malloc(d.n * sizeof(int));
        if (!d.f[k] || !d.h[k]) return -3;
        /* cell.hpp:9-15: pdf = weights; domain.hpp:87-93: fluid handler everywhere */
        for (size_t i = 0; i < d.n; ++i) {
            for (int q = 0; q < Q; ++q) d.f[k][i * Q + q] = d.m.w[q];
            d.h[k][i] = H_FLUID;
        }
    }
    d.collide = 0;
    d.handlers = (handler_t*) calloc((size_t) c->n_boxes + 1, sizeof(handler_t));

    if (c->fluid_mask) {           /* io/vtk.hpp:137-150 */
        const int h = c->n_boxes;  /* one extra NoSlip record */
        d.handlers[h].kind = ORC_NOSLIP;
        size_t i = 0;
        for (int z = 1; z < d.zl + 1; ++z)
            for (int y = 1; y < d.yl + 1; ++y)
                for (int x = 1; x < d.xl + 1; ++x)
                    if (!c->fluid_mask[i++]) {
                        d.h[d.collide][idx(&d, x, y, z)] = h;
                        if (!c->mask_literal) d.h[1 - d.collide][idx(&d, x, y, z)] = h;
                    }
    }
    for (int b = 0; b < c->n_boxes; ++b) {   /* io/scenario.h:183-186: document order */
        const orc_box* box = &c->boxes[b];
        if (!(box->xE >= box->x0 && box->yE >= box->y0 && box->zE >= box->z0)) return -4;
        if (!(box->xE < c->xl + 2 && box->yE < c->yl + 2 && box->zE < c->zl + 2)) return -4;
        d.handlers[b].kind = box->kind;
        memcpy(d.handlers[b].v, box->v, sizeof box->v);
        d.handlers[b].rho = box->rho;
        set_boundary(&d, b, box);
    }
    if (c->f_init) memcpy(d.f[d.collide], c->f_init, d.n * Q * sizeof(double));
    if (c->null_opt) set_nonfluid_cells_nullcollide(&d);

    double seconds = 0.0;
    for (uint64_t t = 1; t <= c->steps; ++t) {
        if (c->periodic) {   /* ghost = wrapped interior, whole Cell assigned */
            for (int z = 0; z < d.zl + 2; ++z)
                for (int y = 0; y < d.yl + 2; ++y)
                    for (int x = 0; x < d.xl + 2; ++x)
                        if (!in_bounds(&d, x, y, z)) {
                            const int wx = wrap(x, d.xl), wy = wrap(y, d.yl), wz = wrap(z, d.zl);
                            memcpy(cell(&d, x, y, z), cell(&d, wx, wy, wz), Q * sizeof(double));
                            d.h[d.collide][idx(&d, x, y, z)] = handler(&d, wx, wy, wz);
                        }
        }
        const double start = omp_get_wtime();   /* src/main.cpp:49-53 */
        stream(&d);
        swap_fields(&d);
        collide(&d);
        if (t > c->untimed) seconds += omp_get_wtime() - start;
    }
    r->seconds = seconds;

    if (r->f) memcpy(r->f, d.f[d.collide], d.n * Q * sizeof(double));
    if (r->kind)
        for (size_t i = 0; i < d.n; ++i) r->kind[i] = (uint8_t) kind_of(&d, d.h[d.collide][i]);
    if (r->rho || r->u) {        /* io/vtk.hpp:62-73 */
        size_t i = 0;
        for (int z = 1; z < d.zl + 1; ++z)
            for (int y = 1; y < d.yl + 1; ++y)
                for (int x = 1; x < d.xl + 1; ++x, ++i) {
                    double u[3];
                    const double* f = cell(&d, x, y, z);
                    const double density = density_of(&d.m, f);
                    velocity_of(&d.m, f, density, u);
                    if (r->rho) r->rho[i] = density;
                    if (r->u) { r->u[3 * i] = u[0]; r->u[3 * i + 1] = u[1]; r->u[3 * i + 2] = u[2]; }
                }
    }
    for (int k = 0; k < 2; ++k) { free(d.f[k]); free(d.h[k]); }
    free(d.handlers);
    return 0;
}

double oracle_density(int Q, const double* f)
{
    model_t m;
    if (make_model(Q, &m)) return NAN;
    return density_of(&m, f);
}
void oracle_velocity(int Q, const double* f, double density, double* u3)
{
    model_t m;
    if (make_model(Q, &m)) return;
    velocity_of(&m, f, density, u3);
}
void oracle_feq(int Q, double density, const double* u3, double* feq)
{
    model_t m;
    if (make_model(Q, &m)) return;
    feq_of(&m, density, u3, feq);
}
void oracle_bgk(int Q, double tau, double* f)
{
    model_t m;
    if (make_model(Q, &m)) return;
    bgk_collide(&m, tau, f);
}
int oracle_model(int Q, double* velocities, double* weights)
{
    model_t m;
    if (make_model(Q, &m)) return -1;
    for (int q = 0; q < Q; ++q) {
        for (int k = 0; k < 3; ++k) velocities[q * 3 + k] = m.c[q][k];
        weights[q] = m.w[q];
    }
    return 0;
}
int oracle_velocity_index(int Q, int u, int v, int w) { return velocity_index(Q, u, v, w); }

/*
 * oracle/oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Common C interface of the two CPU checkers used by tests/, smoke() and the
 * cpu_baseline / --impl reference legs of bench.py:
 *
 *   oracle/liblbm_oracle.so      plain-C restatement of the reference algorithm
 *                                (oracle/lbm_oracle.c), symbol prefix  oracle_
 *   oracle/_ref/libref_lbm.so    the reference's OWN headers compiled from
 *                                /root/reference/include (oracle/ref_driver.cpp),
 *                                symbol prefix  ref_
 *
 * Nothing under lbm_b200/ (the product) may include, link or call this.
 *
 * Cell indexing everywhere is the reference's Domain::idx (domain.hpp:61-64):
 *   idx = x + (xl+2)*y + (xl+2)*(yl+2)*z ,  x,y,z in 0..l+1 (ghost shell included)
 * Populations are exchanged in the reference's array-of-structs order
 *   f[idx*Q + q]   (Cell::pdf, cell.h:14).
 */
#ifndef LBM_ORACLE_H
#define LBM_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* handler kinds; numbering shared with include/lbm_b200.h */
enum {
    ORC_FLUID = 0,      /* BGKCollision           collision.h:64-71      */
    ORC_NOSLIP = 1,     /* NoSlipBoundary         boundary.hpp:15-31     */
    ORC_MOVINGWALL = 2, /* MovingWallBoundary     boundary.hpp:44-68     */
    ORC_FREESLIP = 3,   /* FreeSlipBoundary       boundary.hpp:80-115    */
    ORC_OUTFLOW = 4,    /* OutflowBoundary        boundary.hpp:129-150   */
    ORC_INFLOW = 5,     /* InflowBoundary         boundary.hpp:165-181   */
    ORC_PRESSURE = 6,   /* PressureBoundary       boundary.hpp:195-214   */
    ORC_NULL = 7,       /* NullCollision          collision.h:74-86      */
    ORC_PARALLEL = 8    /* parallel::ParallelBoundary (no-op) parallel.h:11-23 */
};

/* One Domain::setBoundaryCondition call (domain.hpp:175-194): an inclusive
 * index box and the handler (kind + constructor arguments) applied to it.
 * Boxes are applied in array order; later boxes overwrite earlier ones. */
typedef struct {
    int32_t kind;        /* ORC_NOSLIP .. ORC_PRESSURE, ORC_PARALLEL          */
    int32_t _pad;
    double v[3];         /* wall_velocity / inflow_velocity                    */
    double rho;          /* reference_density / input_density                  */
    uint64_t x0, xE, y0, yE, z0, zE;
} orc_box;

typedef struct {
    int32_t Q;           /* 15, 19 or 27                                       */
    int32_t threads;     /* OpenMP threads (>=1)                               */
    uint64_t xl, yl, zl; /* interior lengths                                   */
    double tau;
    int32_t n_boxes;
    int32_t periodic;    /* !=0: before every stream(), every ghost-shell cell */
                         /* is overwritten with its periodically wrapped       */
                         /* interior image through Domain::cell()  */
    const orc_box* boxes;
    const uint8_t* fluid_mask; /* optional xl*yl*zl bytes, x fastest, like the  */
                         /* POINT_DATA of a legacy-VTK mask (io/vtk.hpp:141-150):*/
                         /* 0 => interior cell gets a NoSlipBoundary BEFORE the */
                         /* boxes are applied.  NULL => all interior fluid.     */
    const double* f_init;/* optional (xl+2)(yl+2)(zl+2)*Q AoS values written to */
                         /* the collide field through cell(x,y,z)[q]; NULL =>   */
                         /* reference default (weights, cell.hpp:9-15)          */
    int32_t null_opt;    /* !=0: call set_nonfluid_cells_nullcollide()          */
    int32_t mask_literal;/* !=0: fluid_mask handlers are set on the collide     */
                         /* field only, exactly as io/vtk.hpp:145-146 does (the */
                         /* stream field keeps the fluid handler, so masked     */
                         /* cells flip solid/fluid on every swap()); 0 => set   */
                         /* on both fields like setBoundaryCondition does       */
    uint64_t steps;      /* number of stream(); swap(); collide(); iterations   */
    uint64_t untimed;    /* the first `untimed` of those iterations are warm-up: */
                         /* they run but are left out of orc_result.seconds      */
} orc_case;

typedef struct {
    double* f;           /* out, optional: collide field, AoS, all cells        */
    double* rho;         /* out, optional: xl*yl*zl, z,y,x order (vtk.hpp:62-73)*/
    double* u;           /* out, optional: xl*yl*zl*3, same order               */
    uint8_t* kind;       /* out, optional: handler kind of every cell           */
    double seconds;      /* out: wall time inside stream+swap+collide only      */
                         /*      (the region timed by src/main.cpp:49-53)       */
} orc_result;

/* plain-C restatement */
int oracle_run(const orc_case* c, orc_result* r);
/* single-cell helpers of the restatement (collision.hpp:7-70) */
double oracle_density(int Q, const double* f);
void oracle_velocity(int Q, const double* f, double density, double* u3);
void oracle_feq(int Q, double density, const double* u3, double* feq);
void oracle_bgk(int Q, double tau, double* f);
int oracle_model(int Q, double* velocities /*Q*3*/, double* weights /*Q*/);
int oracle_velocity_index(int Q, int u, int v, int w);

/* the reference's own code (only in oracle/_ref/libref_lbm.so) */
int ref_run(const orc_case* c, orc_result* r);
double ref_density(int Q, const double* f);
void ref_velocity(int Q, const double* f, double density, double* u3);
void ref_feq(int Q, double density, const double* u3, double* feq);
void ref_bgk(int Q, double tau, double* f);
int ref_model(int Q, double* velocities, double* weights);
int ref_velocity_index(int Q, int u, int v, int w);

#ifdef __cplusplus
}
#endif
#endif

/* ode_b200.h -- C ABI of libode_b200.so: the B200 (sm_100a) implementation of
 * ODE.jl's Runge-Kutta / Rosenbrock integration hot path.
 *
 * The reference (SciML/ODE.jl v2.15.0, pure Julia) has no FFI layer; the
 * boundary this ABI replaces is the call of the drivers
 *     oderk_adapt   src/runge_kutta.jl:232-369   (ode21/ode23/ode45_fe/ode45_dp/ode45/ode78, :218-224)
 *     oderk_fixed   src/runge_kutta.jl:176-212   (ode1/ode2_midpoint/ode2_heun/ode4, :165-168)
 *     ode23s        src/ODE.jl:252-361
 * made by the `odeXX(F, y0, tspan; reltol, abstol, minstep, maxstep, initstep,
 * points)` front-ends and by `solve(prob, alg)` (src/common.jl:93-100).
 * A Julia closure cannot run on the device, so `F` crosses the ABI as a
 * registered right-hand-side id plus a parameter vector.
 *
 * Conventions
 *  - plain C, no C++/torch types; every entry point returns 0 on success or a
 *    negative ODE_B200_E_* code and never unwinds; ode_b200_last_error() gives a
 *    thread-local message.
 *  - the caller owns every buffer it passes for the duration of the call; the
 *    library owns device memory and the opaque handles it returns.
 *  - all arithmetic is IEEE binary64 with no fused multiply-add in the solver
 *    arithmetic and the reference's association order; `^` and `log10` are the
 *    deterministic ones of ode_b200_detmath.h.
 *  - there is no CPU fallback: without a usable CUDA device every solve entry
 *    point fails with ODE_B200_E_NO_DEVICE.
 */
#ifndef ODE_B200_H
#define ODE_B200_H

#ifdef __CUDACC_RTC__ /* NVRTC has no <stdint.h>; the run-time RHS compiler includes this header */
typedef int int32_t;
typedef unsigned int uint32_t;
typedef long long int64_t;
typedef unsigned long long uint64_t;
#else
#include <stddef.h>
#include <stdint.h>
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define ODE_B200_ABI_VERSION 1
/* doubles per rank in the exchange buffer of the slab sessions (see ode_b200_slab_*) */
#define ODE_B200_XCHG_DOUBLES 264
/* most slabs (GPUs of one node) that can exchange their records over peer memory (ode_b200_slab_peer_*) */
#define ODE_B200_MAX_PEERS 16
/* bytes of the opaque handle ode_b200_slab_peer_alloc exports for the other processes (a cudaIpcMemHandle_t) */
#define ODE_B200_PEER_HANDLE_BYTES 64
/* limits of a registered large-system stencil (ode_b200_register_stencil) */
#define ODE_B200_STENCIL_MAX_RADIUS 4
#define ODE_B200_STENCIL_MAX_NP 8

/* ---- methods: one per reference tableau / driver ------------------------- */
typedef enum {
  ODE_B200_M_FEULER = 0,   /* ode1           bt_feuler   src/runge_kutta.jl:66  */
  ODE_B200_M_MIDPOINT = 1, /* ode2_midpoint  bt_midpoint :71 */
  ODE_B200_M_HEUN = 2,     /* ode2_heun      bt_heun     :77 */
  ODE_B200_M_RK4 = 3,      /* ode4           bt_rk4      :83 */
  ODE_B200_M_RK21 = 4,     /* ode21          bt_rk21     :93 */
  ODE_B200_M_RK23 = 5,     /* ode23          bt_rk23     :101 */
  ODE_B200_M_RK45_FE = 6,  /* ode45_fe       bt_rk45     :112 */
  ODE_B200_M_DOPRI5 = 7,   /* ode45 = ode45_dp  bt_dopri5 :124 */
  ODE_B200_M_FEH78 = 8,    /* ode78          bt_feh78    :140 */
  ODE_B200_M_ODE23S = 9,   /* ode23s         src/ODE.jl:252 */
  ODE_B200_M_ODE4S_S = 10, /* ode4s = ode4s_s  oderosenbrock with Shampine coefficients, src/ODE.jl:366-438 */
  ODE_B200_M_ODE4S_KR = 11,/* ode4s_kr         oderosenbrock with Kaps-Rentrop coefficients, src/ODE.jl:408-419 */
  ODE_B200_M_ODE4MS = 12,  /* ode4ms           Adams-Bashforth order 4, ode_ms src/ODE.jl:175-212 */
  ODE_B200_M_ODE5MS = 13,  /* ode5ms           ode_ms(F, x0, tspan, 5) src/ODE.jl:213; the order-5 row is computed in floating
                              point the way the reference does it with Polynomials (:183-194) */
  ODE_B200_M_ODE_MS1 = 14, /* ode_ms(F, x0, tspan, 1): the first row of ms_coefficients4 (src/ODE.jl:180-181,440) */
  ODE_B200_M_ODE_MS2 = 15, /* ode_ms(F, x0, tspan, 2) */
  ODE_B200_M_ODE_MS3 = 16, /* ode_ms(F, x0, tspan, 3) */
  ODE_B200_M_CK45 = 17,    /* ode45_ck: Cash-Karp 4(5), order (5,4).  EXTENSION: README.md:57 names the coefficients, the
                              reference source has no such tableau; literature coefficients, driver =
                              oderk_adapt src/runge_kutta.jl:232-369 */
  ODE_B200_M_COUNT = 18
} ode_b200_method;
/* order of an ode_ms method id (1..5), 0 for every other method */
#define ODE_B200_MS_MAX_ORDER 5

/* ---- registered right-hand sides F(t, y; params) ------------------------- */
typedef enum {
  ODE_B200_RHS_LINEAR = 0,    /* dof 1: a*u              params {a}   (README.md:19, test/runtests.jl:54) */
  ODE_B200_RHS_CONST = 1,     /* dof 1: c                params {c}   (test/runtests.jl:40) */
  ODE_B200_RHS_TIMELIN = 2,   /* dof 1: c*t              params {c}   (test/runtests.jl:47) */
  ODE_B200_RHS_ROTATION = 3,  /* dof 2: [-y2, y1]                     (test/runtests.jl:67) */
  ODE_B200_RHS_LORENZ = 4,    /* dof 3: params {sigma, rho, beta}     (examples/Lorenz_Attractor.ipynb) */
  ODE_B200_RHS_VDP = 5,       /* dof 2: [y2, mu*((1-y1*y1)*y2) - y1]  params {mu} */
  ODE_B200_RHS_ROBERTSON = 6, /* dof 3: params {k1, k2, k3}           (test/runtests.jl:81-87) */
  ODE_B200_RHS_KEPLER = 7,    /* dof 4: planar two-body, mu = 1 */
  ODE_B200_RHS_REACTION_DIFFUSION = 8, /* dof N: kappa*((u[i-1]-2u[i])+u[i+1]) + (r*u[i])*(1-u[i]), periodic; params {kappa, r} */
  ODE_B200_RHS_RD_FIELD = 9,  /* dof N: the same reaction-diffusion system written as a general gather "field" (one
                                  kernel per stage instead of the fused stencil kernel); params {kappa, r} */
  ODE_B200_RHS_COUNT = 10
} ode_b200_rhs;
/* ids returned by ode_b200_register_rhs / ode_b200_register_stencil start here */
#define ODE_B200_RHS_USER_BASE 1000
#define ODE_B200_STENCIL_USER_BASE 2000
#define ODE_B200_FIELD_USER_BASE 3000

/* `points` keyword (src/runge_kutta.jl:238,283-290); FINAL is an ensemble-only
 * extension that skips all intermediate output. */
typedef enum {
  ODE_B200_POINTS_ALL = 0,
  ODE_B200_POINTS_SPECIFIED = 1,
  ODE_B200_POINTS_FINAL = 2
} ode_b200_points;

/* per-trajectory status */
typedef enum {
  ODE_B200_ST_OK = 0,
  ODE_B200_ST_MINSTEP = 1,   /* "Warning: dt < minstep.  Stopping." src/runge_kutta.jl:358-360; truncated result */
  ODE_B200_ST_MAXSTEPS = 2,  /* max_steps guard hit (the reference would keep looping) */
  ODE_B200_ST_NONFINITE = 3, /* dt became NaN/Inf (the reference would loop forever) */
  ODE_B200_ST_SINGULAR = 4,  /* ode23s: zero pivot in W = I - h*d*J (Julia raises SingularException) */
  ODE_B200_ST_OVERFLOW = 5,  /* points buffer too small; n_points holds the required count */
  ODE_B200_ST_EXCHANGE = 6   /* domain-decomposed system: a peer's exchange record did not arrive in time */
} ode_b200_status;

/* error codes */
#define ODE_B200_OK 0
#define ODE_B200_E_INVALID (-1)      /* bad argument */
#define ODE_B200_E_ZERO_SPAN (-2)    /* "Zero time span"             src/ODE.jl:88 */
#define ODE_B200_E_INITSTEP (-3)     /* "initstep has wrong sign."   src/runge_kutta.jl:294 */
#define ODE_B200_E_POINTS (-4)       /* "Unrecognized option points" src/runge_kutta.jl:289 */
#define ODE_B200_E_NOT_ADAPTIVE (-5) /* "Can only use this solver with an adaptive RK Butcher table" :249 */
#define ODE_B200_E_NO_DEVICE (-6)    /* no CUDA device / driver */
#define ODE_B200_E_CUDA (-7)         /* CUDA runtime error, see last_error */
#define ODE_B200_E_UNSUPPORTED (-8)  /* rhs/method/dof combination not compiled */
#define ODE_B200_E_NOMEM (-9)

/* Options.  The host shim fills the reference's defaults
 * (reltol 1e-5, abstol 1e-8, minstep span/1e18, maxstep span/2.5, initstep 0,
 * points ALL: src/runge_kutta.jl:233-239, src/ODE.jl:253-259). */
typedef struct {
  int32_t method;    /* ode_b200_method */
  int32_t rhs;       /* ode_b200_rhs */
  int32_t dof;       /* state length per trajectory (N for the large system) */
  int32_t n_params;  /* parameters per trajectory */
  int32_t params_stride; /* 0: one shared parameter vector; else doubles between trajectories */
  int32_t points;    /* ode_b200_points */
  int32_t jacobian;  /* ode23s / ode4s: 0 = forward-difference fdjacobian (reference default, src/ODE.jl:232-243),
                        1 = the registered analytic Jacobian of the rhs (the `jacobian = G(t,y)` keyword) */
  int32_t device;    /* CUDA device ordinal */
  int64_t max_steps; /* guard on step attempts per trajectory; <=0 -> 100000000 */
  double reltol, abstol, minstep, maxstep, initstep;
  int32_t mathmode;  /* 0 = detmath pow/log10 (only mode of the library); the oracle also accepts 1 = libm */
  int32_t host_output; /* ode_b200_solve_ensemble with output buffers in page-locked, device-mapped host memory:
                          0 = stage results in HBM and copy them back after the integration (default);
                          1 = the kernel writes each trajectory's results straight into those buffers when it
                              finishes (no copy-back phase: ~6 % less end-to-end time on one GPU; no gain
                              when many GPUs move data through one host at once -- measured with 8).
                              Pageable buffers are always staged. */
  int32_t n_gpus;    /* multi-GPU fan-out inside the library: 0 or 1 = the one device `device`;
                        N > 1 = devices device .. device+N-1 of this node.  ode_b200_solve_ensemble shards the
                        trajectories (contiguous blocks, one host thread per GPU, results into disjoint slices of the
                        caller's buffers, no collective); ode_b200_solve_large splits the grid into N slabs that exchange
                        their per-attempt records over NVLink peer memory (see ode_b200_solve_large_multi). */
  int32_t fp_mode;   /* the `strict_fp` switch of SURVEY.md section 8b.  ODE_B200_FP_STRICT (0, default): no fused
                        multiply-add, reference association order -- the parity mode, bit-identical to the oracle.
                        ODE_B200_FP_FMA (1): the same kernels compiled with FMA contraction (fewer FP64 instructions,
                        results differ from the reference in the last bits; step counts can differ).  Opt-in, ensemble
                        ode45 family only; anything without an FMA build fails with ODE_B200_E_UNSUPPORTED. */
  int32_t reserved[4];
} ode_b200_opts;
#define ODE_B200_FP_STRICT 0
#define ODE_B200_FP_FMA 1

/* Ensemble I/O.  Host arrays are row-major [trajectory][...]; NULL outputs are
 * skipped.  `tspan` is shared by all trajectories (tspan[0] = start,
 * tspan[n_tspan-1] = end, interior entries are output points). */
typedef struct {
  int64_t n_traj;
  const double* y0;     /* [n_traj][dof] */
  const double* params; /* [n_traj][n_params] or [n_params] when params_stride == 0 */
  const double* tspan;  /* [n_tspan] */
  int32_t n_tspan;
  int32_t pad0;
  double* t_final;      /* [n_traj]  last output time (tout[end]) */
  double* y_final;      /* [n_traj][dof]  yout[end] */
  int32_t* n_accept;    /* [n_traj] */
  int32_t* n_reject;    /* [n_traj] */
  int32_t* n_rhs;       /* [n_traj] RHS evaluations */
  int32_t* status;      /* [n_traj] ode_b200_status */
  /* points output (ALL / SPECIFIED): tout, yout per trajectory, capacity pts_cap points each */
  double* pts_t;        /* [n_traj][pts_cap] */
  double* pts_y;        /* [n_traj][pts_cap][dof] */
  int32_t* pts_n;       /* [n_traj] points written (required count on overflow) */
  int64_t pts_cap;
  /* optional attempt trace of ONE trajectory: rows (t, dt, err, accepted) */
  int64_t trace_traj;   /* trajectory index, -1 = none */
  double* trace;        /* [trace_cap][4] */
  int64_t trace_cap;
  int64_t* trace_n;     /* attempts recorded */
} ode_b200_ensemble_io;

#ifndef __CUDACC_RTC__ /* host entry points: not visible to the run-time (NVRTC) compilation of kernels */
/* ---- library info --------------------------------------------------------- */
int ode_b200_version(void);                 /* ODE_B200_ABI_VERSION */
const char* ode_b200_last_error(void);      /* thread-local, never NULL */
int ode_b200_device_count(void);            /* CUDA devices visible, 0 if none */
int ode_b200_rhs_id(const char* name);      /* "lorenz" -> id, <0 unknown */
int ode_b200_method_id(const char* name);   /* "ode45" -> id, <0 unknown */
int ode_b200_rhs_dof(int rhs);              /* fixed dof of a registered rhs, 0 = variable */
int ode_b200_rhs_nparams(int rhs);
int ode_b200_method_stages(int method);     /* S of the tableau (src/runge_kutta.jl:66-157), 3 for ode23s */
int ode_b200_method_is_adaptive(int method);/* isadaptive, src/runge_kutta.jl:58 */
int ode_b200_method_is_fsal(int method);    /* isFSAL, src/runge_kutta.jl:62 */
int ode_b200_method_ms_order(int method);   /* order of ode_ms for ode4ms / ode5ms / ode_ms1..3, 0 for other methods */
/* b[order][j] of ode_ms (src/ODE.jl:180-194): rows 1-4 are ms_coefficients4 (:440-443), row 5 the reference's
 * floating-point evaluation; order, j 1-based, j <= order; NaN outside */
double ode_b200_ms_coefficient(int order, int j);
/* Tableau entry as the reference converts it (Rational -> Float64, src/ODE.jl:163).
 * which: 0 = a[i][j], 1 = b[i][j] (i = 0 stepping row, 1 error row), 2 = c[i]. */
double ode_b200_tableau(int method, int which, int i, int j);

/* ---- right-hand sides registered at run time ------------------------------- */
/* The reference accepts any Julia function F(t, y).  The equivalent here: register the BODY of the device
 * functor as CUDA C++ source.  `eval_body` is the body of
 *     void eval(double t, const double (&y)[dof], double (&f)[dof], const double (&p)[max(n_params,1)])
 * and `jac_body` (optional, NULL = none) the body of
 *     void jac (double t, const double (&y)[dof], double (&J)[dof][dof], const double (&p)[...])
 * (row-major dF/dy, used when opts.jacobian = 1).  On first use the library compiles its own kernel templates
 * for this functor with NVRTC (same arithmetic rules: no FMA contraction) and caches the result.
 * Returns the rhs id (>= ODE_B200_RHS_USER_BASE) to put in opts.rhs, or a negative error code.
 * Available for the ensemble / single-system entry points (every method). */
int ode_b200_register_rhs(const char* name, int32_t dof, int32_t n_params, const char* eval_body, const char* jac_body);
/* Same for the large-system regime: a stencil of radius 1..4 on a periodic 1-D grid,
 *     F_i(t, u) = point(w, p, t, i, n),   w[k] = u[i - radius + k]  (k = 0 .. 2*radius),
 * p the parameter vector (n_params <= 8), t the stage time, i the global grid index, n the ring length.
 * `point_body` is the body of
 *     double point(const double (&w)[2*radius+1], const double (&p)[8], double t, long long i, long long n)
 * (must `return` the value; a 3-point stencil may also use the names um, u, up and p0, p1).  Returns an rhs id
 * >= ODE_B200_STENCIL_USER_BASE for ode_b200_solve_large* and the slab sessions. */
int ode_b200_register_stencil(const char* name, int32_t radius, int32_t n_params, const char* point_body);
/* A GENERAL right-hand side for one large system: F_i(t, y) may read any component of y (non-local couplings,
 * networks, N-body sums, 2-D grids flattened by the caller).  `rhs_body` is the body of
 *     double rhs(long long i, long long n, const double* y, const double (&p)[8], double t)
 * returning dy_i/dt (n_params <= 8).  Such a system runs one kernel per Runge-Kutta stage (the stage argument must be
 * complete in memory before any component of F can be evaluated) on one GPU, through ode_b200_solve_large* and a
 * single-slab session (world = 1).  Returns an rhs id >= ODE_B200_FIELD_USER_BASE. */
int ode_b200_register_field(const char* name, int32_t n_params, const char* rhs_body);

/* ---- ensembles of small systems (one trajectory per thread) -------------- */
/* Replaces n_traj independent calls of oderk_adapt / oderk_fixed / ode23s.
 * Host buffers; every transfer happens inside the (blocking) call.  For large adaptive-RK ensembles the transfers
 * run beside the one persistent kernel instead of around it: y0 in page-locked memory (one parameter vector) is read
 * by the hinit pre-pass over PCIe while the persistent kernel already integrates; outputs in page-locked memory are
 * written by the kernel as trajectories end (opts.host_output = 1), staged outputs are copied back chunk by chunk
 * while it runs.  Pageable buffers work too (uploaded first / copied through the driver's staging). */
int ode_b200_solve_ensemble(const ode_b200_opts* opts, const ode_b200_ensemble_io* io);
/* Same, but every pointer in `io` is a device pointer on opts->device (tspan
 * included); work is enqueued on `stream` (a cudaStream_t, NULL = default
 * stream) and the call returns without waiting for the kernels.  Scratch memory is
 * taken per call from the device's stream-ordered pool, so several calls may be in
 * flight on different streams (from one thread or many). */
int ode_b200_solve_ensemble_dev(const ode_b200_opts* opts, const ode_b200_ensemble_io* io, void* stream);
/* Milliseconds the last ensemble kernel of this thread took (CUDA events on its stream). */
double ode_b200_last_kernel_ms(void);
/* Number of kernel launches issued by this thread since the last reset. */
int64_t ode_b200_launch_count(int reset);

/* ---- one system, reference return contract (tout, yout) ------------------ */
typedef struct ode_b200_result ode_b200_result;
/* Replaces one call of odeXX(F, y0, tspan; ...): result holds tout/yout with
 * the reference's length rules (src/runge_kutta.jl:321-340, src/ODE.jl:329-345). */
int ode_b200_solve_single(const ode_b200_opts* opts, const double* y0, const double* params,
                          const double* tspan, int32_t n_tspan, ode_b200_result** out);
int64_t ode_b200_result_len(const ode_b200_result* r);
int32_t ode_b200_result_dof(const ode_b200_result* r);
int32_t ode_b200_result_status(const ode_b200_result* r);
int64_t ode_b200_result_counts(const ode_b200_result* r, int which); /* 0 accepted, 1 rejected, 2 rhs calls */
int ode_b200_result_copy(const ode_b200_result* r, double* tout, double* yout /*[len][dof]*/);
void ode_b200_result_free(ode_b200_result* r);

/* ---- one large method-of-lines system (fused stage kernels, HBM-bound) ---- */
typedef struct {
  int64_t n_accept, n_reject, n_rhs;
  int32_t status;
  int32_t pad;
  double t_final;
  double kernel_ms; /* device time of the whole integration loop */
} ode_b200_large_stats;

/* Replaces one call of oderk_adapt on a dof = N state with the registered
 * stencil RHS.  Host buffers in/out, one GPU.  trace (optional): rows
 * (t, dt, err, accepted) per attempt. */
int ode_b200_solve_large(const ode_b200_opts* opts, int64_t n, const double* y0, const double* params,
                         double t0, double tend, double* y_out, ode_b200_large_stats* stats,
                         double* trace, int64_t trace_cap, int64_t* trace_n);

/* The same call fanned out over n_gpus devices of this process (devices[] = CUDA ordinals, NULL = opts->device ..
 * opts->device + n_gpus - 1; n_gpus <= ODE_B200_MAX_PEERS, every pair needs peer access, i.e. NVLink/NVSwitch or
 * PCIe P2P).  The reference has one Vector in one process (oderk_adapt, src/runge_kutta.jl:232-369); here the
 * periodic grid is cut into n_gpus contiguous slabs (the first n % n_gpus one point longer).  Per step attempt every
 * device runs one kernel whose last thread block writes the slab's record (exactly accumulated error partial, NaN
 * flag, slab-edge values) into every peer's HBM, waits for theirs and runs the step controller on the gathered
 * data, so all devices take the same decision and the result is the single-GPU result bit for bit.  y0 / y_out
 * are host arrays of n doubles (each device reads and writes its own slice).  ode_b200_solve_large with
 * opts->n_gpus > 1 calls this. */
int ode_b200_solve_large_multi(const ode_b200_opts* opts, int32_t n_gpus, const int32_t* devices, int64_t n,
                               const double* y0, const double* params, double t0, double tend, double* y_out,
                               ode_b200_large_stats* stats, double* trace, int64_t trace_cap, int64_t* trace_n);

/* oderk_fixed (src/runge_kutta.jl:176-212: ode1, ode2_midpoint, ode2_heun, ode4) and ode_ms (src/ODE.jl:175-213: ode4ms, ode5ms, orders 1-3) on a
 * dof = N state with a stencil or field right-hand side: one step per tspan interval, one kernel per stage.
 * pts_y (optional) is [n_tspan][n] and
 * receives one state per tspan entry like the reference's yout; y_final (optional) the last one. */
int ode_b200_solve_large_fixed(const ode_b200_opts* opts, int64_t n, const double* y0, const double* params,
                               const double* tspan, int32_t n_tspan, double* pts_y, double* y_final,
                               ode_b200_large_stats* stats);

/* The host-buffer entry points keep device memory of the calling thread's last call for reuse (the ensemble
 * arena, the large-system session); this releases it. */
void ode_b200_trim(void);

/* Page-locked, device-mapped host memory for the caller's buffers (cudaHostAlloc).  The reference has no
 * counterpart (Julia's GC owns its arrays); a binding uses these for arrays it hands to the entry points
 * above repeatedly: ode_b200_solve_ensemble lets the kernel write results straight into output buffers that
 * live in such memory (no D2H phase after the integration), and copies from/to them at full PCIe rate.
 * Buffers from any other allocator (malloc, a GC heap) work everywhere, staged through HBM. */
void* ode_b200_host_alloc(size_t bytes);
void ode_b200_host_free(void* p);

/* Same with the reference's points = :specified contract (src/runge_kutta.jl:321-326): pts_y is
 * [n_tspan][n]; row 0 is y0, the other rows are 3rd-order Hermite values (hermite_interp!, :479-492)
 * at tspan[i], the last row at tspan[end]. */
int ode_b200_solve_large_points(const ode_b200_opts* opts, int64_t n, const double* y0, const double* params,
                                const double* tspan, int32_t n_tspan, double* pts_y, ode_b200_large_stats* stats,
                                double* trace, int64_t trace_cap, int64_t* trace_n);

/* Same with points = :all (src/runge_kutta.jl:327-340): every accepted step's state, preceded by the Hermite
 * values at interior tspan points, capped at cap_rows rows: pts_t[cap_rows], pts_y[cap_rows][n].  *n_rows is
 * the number of rows the reference would return; if it exceeds cap_rows, stats->status is
 * ODE_B200_ST_OVERFLOW and the caller reruns with *n_rows rows (the integration is deterministic). */
int ode_b200_solve_large_all(const ode_b200_opts* opts, int64_t n, const double* y0, const double* params,
                             const double* tspan, int32_t n_tspan, int64_t cap_rows, double* pts_t, double* pts_y,
                             int64_t* n_rows, ode_b200_large_stats* stats, double* trace, int64_t trace_cap,
                             int64_t* trace_n);

/* Slab session for the domain-decomposed path: every rank owns n_local
 * contiguous points of an n_global periodic grid plus ghost cells; between the
 * two halves of an attempt the host exchanges `xchg_send` (ODE_B200_XCHG_DOUBLES
 * doubles per rank) with an all-gather into `xchg_recv` ([world][XCHG]).
 * All pointers are device pointers; work is enqueued on `stream`. */
typedef struct ode_b200_slab ode_b200_slab;
int ode_b200_slab_create(const ode_b200_opts* opts, int64_t n_local, int64_t n_global, int32_t rank,
                         int32_t world, const double* params, double t0, double tend, void* stream,
                         ode_b200_slab** out);
int ode_b200_slab_set_state(ode_b200_slab* s, const double* y_local_dev); /* [n_local] device */
/* re-arm the session for another integration over [t0, tend] with the same options
 * (the reference is stateless: a restart is a new call); follow with set_state */
int ode_b200_slab_reset(ode_b200_slab* s, double t0, double tend);
/* points = :specified output: tspan (host, n_tspan entries) and a device buffer [n_tspan][n_local];
 * call before set_state (which writes row 0 = y0).  NULL disables. */
int ode_b200_slab_set_output(ode_b200_slab* s, const double* tspan_host, int32_t n_tspan, double* out_dev);
/* points = :all output: rows (t, y) for every accepted step plus interior tspan points, capped at cap_rows;
 * out_dev [cap_rows][n_local], out_times_dev [cap_rows]; ctl_out[12] of poll is the row count needed. */
int ode_b200_slab_set_output_all(ode_b200_slab* s, const double* tspan_host, int32_t n_tspan, double* out_dev,
                                 double* out_times_dev, int64_t cap_rows);
/* global grid index of this slab's first point (what a stencil's `i` argument counts from).  Default:
 * contiguous slabs in rank order, the first n_global % world ranks one point longer. */
int ode_b200_slab_set_offset(ode_b200_slab* s, int64_t first_global_index);
/* Single slab (world = 1, stencil right-hand side): let the last thread block of the attempt kernel run the step
 * controller itself.  `attempt` then needs no exchange after it and `control` only produces the requested output
 * rows -- one copy and one dependent launch less per attempt.  Off by default (the decomposed path needs the
 * all-gather between the two). */
int ode_b200_slab_set_fused(ode_b200_slab* s, int32_t on);
/* Exchange over peer memory instead of a caller-side all-gather: the slabs of one node (one process with several
 * devices, or one process per device) write their records straight into each other's HBM from inside the attempt
 * kernel (NVLink peer stores + a flag per record), and the kernel's last thread block goes on to run the step
 * controller -- no collective library call, no host round trip per attempt.
 *   1. every rank: ode_b200_slab_peer_alloc -> its receive block; `ipc_handle_out` (ODE_B200_PEER_HANDLE_BYTES bytes,
 *      may be NULL) is what the OTHER PROCESSES need to map it, `local_ptr_out` (may be NULL) what other devices of
 *      THIS process need (after cudaDeviceEnablePeerAccess);
 *   2. the caller passes the handles / pointers around (once per session), all ranks reach a barrier;
 *   3. every rank: ode_b200_slab_peer_connect with either `ipc_handles` [world][ODE_B200_PEER_HANDLE_BYTES] or
 *      `ptrs` [world] (entry `rank` is ignored); then another barrier before the first ode_b200_slab_run.
 * A peer that stops answering ends the integration with ODE_B200_ST_EXCHANGE after ODE_B200_PEER_TIMEOUT_MS
 * (environment, default 5000). */
int ode_b200_slab_peer_alloc(ode_b200_slab* s, void* ipc_handle_out, void** local_ptr_out);
int ode_b200_slab_peer_connect(ode_b200_slab* s, const void* ipc_handles, void* const* ptrs);
/* The whole integration of this rank's slab in one call: hinit, then batches of attempts replayed from a CUDA graph
 * until the device-side done flag is set.  For a single slab (world = 1) or slabs connected over peer memory; all
 * ranks call it at the same time.  ctl_out[16] as ode_b200_slab_poll; kernel_ms (optional) = device time. */
int ode_b200_slab_run(ode_b200_slab* s, double* ctl_out /*[16]*/, double* kernel_ms);
/* optional attempt trace on the device: rows (t, dt, err, accepted) */
int ode_b200_slab_set_trace(ode_b200_slab* s, double* trace_dev, int64_t trace_cap);
int ode_b200_slab_get_state(ode_b200_slab* s, double* y_local_dev);
/* hinit (src/ODE.jl:85-112) in phases 0..3; phases 0, 1 and 2 are each followed by an exchange
 * (0: y edges; 1: max|y|, max|f0|, f0 edges; 2: max|f1-f0|; 3: dt) */
int ode_b200_slab_init_phase(ode_b200_slab* s, int phase, double* xchg_send, const double* xchg_recv);
/* one attempt: stages + local error partials -> xchg_send */
int ode_b200_slab_attempt(ode_b200_slab* s, double* xchg_send);
/* controller: reduce gathered partials, accept/reject, new dt, ghost refresh */
int ode_b200_slab_control(ode_b200_slab* s, const double* xchg_recv);
/* copy the control block to the host (synchronises the stream): ctl_out[12] =
 * {done, status, n_accept, n_reject, t, dt, err, t_final, n_rhs, attempts, trace_n, ghost width H,
 *  points=:all rows, 0, 0, 0} */
int ode_b200_slab_poll(ode_b200_slab* s, double* ctl_out /*[16]*/);
void ode_b200_slab_free(ode_b200_slab* s);

/* ---- measurement hooks ---------------------------------------------------- */
/* Register-resident DFMA/DADD/DMUL chain microbenchmark: returns FP64
 * instructions per second of the given kind (0 = DFMA, 1 = DADD+DMUL mix) on
 * opts device, <0 on error.  Used as the ensemble roofline denominator
 * (SURVEY.md F10). */
double ode_b200_fp64_peak(int device, int kind);
/* Evaluates the device build of ode_b200_detmath.h: out_pow[i] = dm_pow(x[i], y[i]),
 * out_log10[i] = dm_log10(x[i]) (host pointers).  Lets a caller verify that the
 * GPU produces the same bits as the host build of the same header. */
int ode_b200_selftest_detmath(int device, const double* x, const double* y, double* out_pow, double* out_log10,
                              int64_t n);
/* Compares the shared-denominator division of the kernels (recip_prepare/div_by, csrc/common.cuh) with the
 * compiler's IEEE a / b on about n_samples pseudo-random operand pairs (all exponent ranges, NaN, Inf, zeros,
 * denormals, hard mantissas).  Returns the number of pairs whose bits differ (0 expected), <0 on error. */
int64_t ode_b200_selftest_div(int device, int64_t n_samples, uint64_t seed);

#endif /* !__CUDACC_RTC__ */

#ifdef __cplusplus
}
#endif
#endif /* ODE_B200_H */

// large.cu -- host side of the large-system regime: slab sessions, the device
// control kernel, hinit in phases, and the single-GPU convenience call.
// Kernels: large.cuh (fused attempt) + the small init/control kernels below.
#include <cuda_runtime.h>

#include <cmath>
#include <cstring>
#include <new>
#include <vector>

#include "host_util.cuh"
#include "field.cuh"
#include "rtc.cuh"

namespace odeb200 {

__global__ void large_control_kernel(const ControlArgs C) { large_control_body(C); }

// hermite_interp! (src/runge_kutta.jl:479-492) for the tspan rows the control kernel selected.
__global__ void __launch_bounds__(256) large_hermite_kernel(double* const Y0, double* const Y1, double* const K0,
                                                            double* const K1, const LargeCtl* ctl, const double* tq,
                                                            double* out, long long n, int H, long long out_cap,
                                                            int all_mode) {
  const int lo = ctl->out_lo, hi = ctl->out_hi;
  const int end_row = all_mode ? ctl->out_end : 0;
  if (hi <= lo && !end_row) return;
  const int par = ctl->out_par;
  const double* y0 = (par ? Y1 : Y0) + H;
  const double* y1 = (par ? Y0 : Y1) + H;
  const double* f0 = (par ? K1 : K0) + H;
  const double* f1 = (par ? K0 : K1) + H;
  const double t = ctl->out_t, dt = ctl->out_dt;
  const long long row0 = all_mode ? ctl->out_row0 : (long long)lo;
  for (int q = lo; q < hi; q++) {
    const long long row = row0 + (q - lo);
    if (all_mode && row >= out_cap) continue;
    const double theta = (tq[q] - t) / dt;  // :486
    double* o = out + row * n;
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
      const double a = y0[i], b = y1[i];
      o[i] = ((1.0 - theta) * a + theta * b) +
             (theta * (theta - 1.0)) *
                 (((1.0 - 2.0 * theta) * (b - a) + ((theta - 1.0) * dt) * f0[i]) + (theta * dt) * f1[i]);  // :488-489
    }
  }
  if (end_row) {  // index_or_push!(ys, iter, deepcopy(ytrial)) :337
    const long long row = row0 + (hi - lo);
    if (row < out_cap) {
      double* o = out + row * n;
      for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += (long long)gridDim.x * 256) o[i] = y1[i];
    }
  }
}

// ---- hinit (src/ODE.jl:85-112) in phases ------------------------------------------
// pack the H edge values of an interior array into slot `slot` (0: A, 1: B)
// (H can exceed the 32 threads of the launch: ode78 with a radius-3/4 stencil has H = 40 / 52)
__global__ void pack_edges_kernel(const double* a, double* xsend, long long n, int H, int slot) {
  for (int j = threadIdx.x; j < H; j += blockDim.x) {
    xsend[XS_EDGE + 2 * H * slot + j] = a[H + j];
    xsend[XS_EDGE + 2 * H * slot + H + j] = a[H + (n - H) + j];
  }
}
__global__ void large_ghost_kernel(double* a, const double* xrecv, long long n, int rank, int world, int H, int slot) {
  const int lr = (rank + world - 1) % world, rr = (rank + 1) % world;
  const double* xl = xrecv + (long long)lr * XCHG + XS_EDGE + 2 * H * slot;
  const double* xr = xrecv + (long long)rr * XCHG + XS_EDGE + 2 * H * slot;
  for (int j = threadIdx.x; j < H; j += blockDim.x) {
    a[j] = xl[H + j];
    a[n + H + j] = xr[j];
  }
}

// The all-gather of the slabs' records as its own launch: the hinit phases, and the attempts when the step
// controller is not fused into the attempt kernel (ODE_B200_PEER_FUSED=0).  Leaves a copy of the gathered records
// where the kernels launched after it expect them (P.gathered = the session's xr buffer).
__global__ void __launch_bounds__(128) peer_exchange_kernel(const PeerXchg P, const double* xsend, int len, LargeCtl* ctl) {
  if (ctl->done) return;
  const double* g = peer_all_gather(P, xsend, len);
  if (!g) {
    if (threadIdx.x == 0) {
      ctl->status = ODE_B200_ST_EXCHANGE;
      ctl->done = 1;
    }
    return;
  }
  for (int j = threadIdx.x; j < P.world * XCHG; j += blockDim.x) P.gathered[j] = g[j];
}

__global__ void reduce_max_kernel(const double* part, int nblocks, int ncol, double* xsend) {
  // one warp; NaN-propagating max of each column
  for (int c = 0; c < ncol; c++) {
    double v = 0.0;
    for (int b = threadIdx.x; b < nblocks; b += 32) v = normInf_step(v, part[b * 2 + c]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = normInf_step(v, __shfl_down_sync(0xffffffffu, v, o));
    if (threadIdx.x == 0) xsend[XS_AUX0 + c] = v;
  }
}
__global__ void __launch_bounds__(256) field_hinit_x1_kernel(const double* __restrict__ y, const double* __restrict__ f0,
                                                             const LargeCtl* ctl, double* x1, long long n) {
  const double th = ctl->tdir * ctl->h0;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += (long long)gridDim.x * 256) x1[i] = y[i] + th * f0[i];  // :100
}
// hinit scalar parts: stage 0 after the (max|y|, max|f0|) exchange, stage 1 after max|f1-f0|
__global__ void hinit_control_kernel(LargeCtl* ctl, const double* xrecv, int world, int stage, int order) {
  if (threadIdx.x != 0) return;
  if (stage == 0) {
    double n0 = 0.0, n1 = 0.0;
    for (int r = 0; r < world; r++) {
      n0 = normInf_step(n0, xrecv[(long long)r * XCHG + XS_AUX0]);
      n1 = normInf_step(n1, xrecv[(long long)r * XCHG + XS_AUX1]);
    }
    const double tau = jl_max(ctl->reltol * n0, ctl->abstol);  // :89
    const double d0 = n0 / tau, d1 = n1 / tau;                 // :90,:92
    ctl->tau = tau;
    ctl->d0 = d0;
    ctl->d1 = d1;
    ctl->h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * (d0 / d1);  // :93-97
  } else {
    double n2 = 0.0;
    for (int r = 0; r < world; r++) n2 = normInf_step(n2, xrecv[(long long)r * XCHG + XS_AUX0]);
    const double h0 = ctl->h0, tdir = ctl->tdir;
    const double d2 = n2 / (ctl->tau * h0);  // :103
    const double dm = jl_max(ctl->d1, d2);
    double h1;
    if (dm <= 1e-15)
      h1 = jl_max(1e-6, 1e-3 * h0);
    else
      h1 = dm_pow(10.0, -(2.0 + dm_log10(dm)) / (double)(order + 1));  // :107-108
    double dt = tdir * jl_min(jl_min(100.0 * h0, h1), tdir * (ctl->tend - ctl->tstart));  // :111
    if (ctl->initstep != 0.0) dt = ctl->initstep;  // src/runge_kutta.jl:293
    ctl->dt = dt;
    ctl->t = ctl->tstart;
    ctl->islast = (fabs(ctl->t + dt - ctl->tend) <= jl_eps(ctl->tend)) ? 1 : 0;  // :302
    ctl->timeout = 0;
    ctl->nrhs = 2;
    ctl->t_final = ctl->tstart;
    ctl->out_iter = 2;  // :286,:304
    ctl->out_lo = ctl->out_hi = 0;
    ctl->out_end = 0;
    ctl->out_rows = 1;  // ys[1] = y0, tspan = [tstart] (:280,:285)
    ctl->out_row0 = 0;
  }
}

// Kernel handles are `const void*`: a __global__ function pointer for the compiled-in stencil, a
// cudaKernel_t for a stencil registered at run time (rtc.cu).
struct MethodInfo {
  const void *attempt, *attempt_literal, *hinit_f0, *hinit_f1;
  int H, VAL, order, S, fsal;
  // general gather right-hand sides (field.cuh): one kernel per stage instead of `attempt`
  bool field = false;
  const void* f_first = nullptr;
  const void* f_stage[FIELD_MAX_STAGES + 1] = {};  // [s], reference stage numbering 2 .. S
  const void* f_f1 = nullptr;
};
static int persistent_grid(const void* k, int device, long long tiles) {
  int sms = 0, per_sm = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, LARGE_THREADS, 0) != cudaSuccess || per_sm < 1) {
    cudaGetLastError();
    per_sm = 4;
  }
  long long g = (long long)sms * per_sm;  // one wave of resident CTAs, each loops over its tiles
  if (g > tiles) g = tiles;
  return (int)(g < 1 ? 1 : g);
}
// Tile geometry of (tableau, stencil radius): the same arithmetic as LargeGeom, for radii known only at run
// time (registered stencils).
template <class Tab>
static void geometry(MethodInfo* m, int radius) {
  constexpr bool fsal = Tab::fsal();
  const int g = radius * (Tab::S - 1 + (fsal ? 0 : 1));
  m->H = (g + 3) / 4 * 4;
  m->VAL = LARGE_THREADS * LARGE_P - 2 * m->H;
  m->order = Tab::ORDER;
  m->S = Tab::S;
  m->fsal = fsal ? 1 : 0;
}
template <class Tab>
static MethodInfo method_info(int radius) {
  MethodInfo m;
  geometry<Tab>(&m, radius);
  static_assert(LargeGeom<Tab, RhsReactionDiffusion>::H == (Tab::S - 1 + (Tab::fsal() ? 0 : 1) + 3) / 4 * 4, "geometry");
  m.attempt = (const void*)large_attempt_kernel<Tab, RhsReactionDiffusion, false>;
  m.attempt_literal = (const void*)large_attempt_kernel<Tab, RhsReactionDiffusion, true>;
  m.hinit_f0 = (const void*)hinit_f0_kernel<RhsReactionDiffusion>;
  m.hinit_f1 = (const void*)hinit_f1_kernel<RhsReactionDiffusion>;
  return m;
}
template <class Tab, class Rhs, int s>
struct FillFieldStages {
  static void run(MethodInfo* m) {
    m->f_stage[s] = (const void*)field_stage_kernel<Tab, Rhs, s>;
    FillFieldStages<Tab, Rhs, s - 1>::run(m);
  }
};
template <class Tab, class Rhs>
struct FillFieldStages<Tab, Rhs, 1> {
  static void run(MethodInfo*) {}
};
template <class Tab>
static MethodInfo field_method_info() {
  MethodInfo m;
  geometry<Tab>(&m, 0);  // no ghost cells
  m.attempt = m.attempt_literal = nullptr;
  m.field = true;
  m.f_first = (const void*)field_first_kernel<Tab>;
  FillFieldStages<Tab, FieldReactionDiffusion, Tab::S>::run(&m);
  m.f_f1 = (const void*)field_f1_kernel<FieldReactionDiffusion>;
  m.hinit_f0 = (const void*)field_hinit_f0_kernel<FieldReactionDiffusion>;
  m.hinit_f1 = (const void*)field_hinit_f1_kernel<FieldReactionDiffusion>;
  return m;
}
static bool pick_field_method(int method, MethodInfo* out) {
  switch (method) {
    case ODE_B200_M_RK21: *out = field_method_info<TabRK21>(); return true;
    case ODE_B200_M_RK23: *out = field_method_info<TabRK23>(); return true;
    case ODE_B200_M_RK45_FE: *out = field_method_info<TabRK45>(); return true;
    case ODE_B200_M_DOPRI5: *out = field_method_info<TabDopri5>(); return true;
    case ODE_B200_M_FEH78: *out = field_method_info<TabFeh78>(); return true;
    case ODE_B200_M_CK45: *out = field_method_info<TabCK45>(); return true;
    default: return false;
  }
}
static bool pick_method(int method, int radius, MethodInfo* out) {
  switch (method) {
    case ODE_B200_M_RK21: *out = method_info<TabRK21>(radius); return true;
    case ODE_B200_M_RK23: *out = method_info<TabRK23>(radius); return true;
    case ODE_B200_M_RK45_FE: *out = method_info<TabRK45>(radius); return true;
    case ODE_B200_M_DOPRI5: *out = method_info<TabDopri5>(radius); return true;
    case ODE_B200_M_FEH78: *out = method_info<TabFeh78>(radius); return true;
    case ODE_B200_M_CK45: *out = method_info<TabCK45>(radius); return true;
    default: return false;
  }
}


template <class Tab, class Rhs, int s>
struct FillFixedStages {
  static void run(const void** k) {
    k[s - 1] = (const void*)field_fixed_stage_kernel<Tab, Rhs, s>;
    FillFixedStages<Tab, Rhs, s - 1>::run(k);
  }
};
template <class Tab, class Rhs>
struct FillFixedStages<Tab, Rhs, 0> {
  static void run(const void**) {}
};
// stage kernels of a fixed-step tableau for the built-in reaction-diffusion system; returns the stage count
static int builtin_fixed_kernels(int method, const void** k) {
  switch (method) {
    case ODE_B200_M_FEULER: FillFixedStages<TabFEuler, FieldReactionDiffusion, TabFEuler::S>::run(k); return TabFEuler::S;
    case ODE_B200_M_MIDPOINT: FillFixedStages<TabMidpoint, FieldReactionDiffusion, TabMidpoint::S>::run(k); return TabMidpoint::S;
    case ODE_B200_M_HEUN: FillFixedStages<TabHeun, FieldReactionDiffusion, TabHeun::S>::run(k); return TabHeun::S;
    case ODE_B200_M_RK4: FillFixedStages<TabRK4, FieldReactionDiffusion, TabRK4::S>::run(k); return TabRK4::S;
    default: return 0;
  }
}
static int fixed_stage_count(int method) {
  switch (method) {
    case ODE_B200_M_FEULER: return TabFEuler::S;
    case ODE_B200_M_MIDPOINT: return TabMidpoint::S;
    case ODE_B200_M_HEUN: return TabHeun::S;
    case ODE_B200_M_RK4: return TabRK4::S;
    default: return 0;
  }
}

}  // namespace odeb200

using namespace odeb200;

struct ode_b200_slab {
  int device = 0, rank = 0, world = 1;
  long long n = 0, n_global = 0, tiles = 0;
  int grid = 1, grid_literal = 1;
  MethodInfo mi;
  cudaStream_t stream = nullptr;
  double *Y[2] = {nullptr, nullptr}, *K[2] = {nullptr, nullptr};
  LargeCtl* ctl = nullptr;
  double *part = nullptr, *ipart = nullptr;
  double* trace = nullptr;
  long long trace_cap = 0;
  double* tq = nullptr;   // device copy of tspan for points = :specified
  int n_tq = 0;
  double* out = nullptr;  // [n_tq][n] (:specified) or [out_cap][n] (:all) device, caller owned
  double* out_times = nullptr;  // [out_cap] device (:all only)
  long long out_cap = 0;
  StencilCtx sc;  // parameters, global index of the first interior point, ring length
  int np = 2;     // parameters the stencil takes
  double* KS[FIELD_MAX_STAGES] = {};  // field path: middle stages
  double* T[2] = {nullptr, nullptr};  // field path: stage-argument ping-pong
  int field_grid = 1;
  bool fused = false;  // stencil path: the attempt kernel's last CTA runs the step controller (single slab, or peers)
  int init_blocks = 0;
  LargeCtl host_ctl;
  LargeCtl* pin_ctl = nullptr;                // 2 page-locked copies of the control block: slabs_run's pipelined polls
  cudaEvent_t ev_poll[2] = {nullptr, nullptr};
  // exchange buffers owned by the session (ode_b200_slab_run; the step-wise API takes the caller's)
  double *xs = nullptr, *xr = nullptr;
  // exchange over peer memory (ode_b200_slab_peer_*)
  bool peer_on = false;
  double* peer_block = nullptr;              // this rank's receive block (cudaMalloc: exportable with CUDA IPC)
  unsigned long long* peer_epoch = nullptr;  // inside peer_block, after the flags
  void* peer_ipc[PEER_MAX] = {};             // blocks opened with cudaIpcOpenMemHandle (closed on free)
  PeerXchg px;
};

struct LargeCache {
  ode_b200_slab* s = nullptr;
  cudaStream_t st = nullptr;
  double* dy = nullptr;
  int device = -1, method = -1, rhs = -1;
  long long n = 0;
};
static thread_local LargeCache g_large_cache;
namespace odeb200 {
void arena_release();  // capi.cu
}
static void large_cache_release();

static int slab_check(ode_b200_slab* s) {
  if (!s) return set_err(ODE_B200_E_INVALID, "null slab");
  ODEB200_CUDA(cudaSetDevice(s->device));
  return 0;
}

extern "C" {

int ode_b200_slab_create(const ode_b200_opts* o, int64_t n_local, int64_t n_global, int32_t rank, int32_t world,
                         const double* params, double t0, double tend, void* stream, ode_b200_slab** out) {
  if (!o || !out) return set_err(ODE_B200_E_INVALID, "null argument");
  *out = nullptr;
  int cnt = 0;
  if (cudaGetDeviceCount(&cnt) != cudaSuccess || cnt <= 0) {
    cudaGetLastError();
    return set_err(ODE_B200_E_NO_DEVICE, "no CUDA device available; libode_b200 has no CPU fallback");
  }
  if (o->device < 0 || o->device >= cnt) return set_err(ODE_B200_E_INVALID, "device %d out of range", o->device);
  if (o->fp_mode != ODE_B200_FP_STRICT)
    return set_err(ODE_B200_E_UNSUPPORTED, "the large-system kernels are built in strict mode only (fp_mode = 0)");
  const bool user_field = o->rhs >= ODE_B200_FIELD_USER_BASE;
  const bool field = user_field || o->rhs == ODE_B200_RHS_RD_FIELD;
  const bool user_stencil = !field && o->rhs >= ODE_B200_STENCIL_USER_BASE;
  if (o->rhs != ODE_B200_RHS_REACTION_DIFFUSION && !user_stencil && !field)
    return set_err(ODE_B200_E_UNSUPPORTED, "the large-system path takes a stencil RHS (reaction_diffusion or a registered stencil) "
                                           "or a registered field");
  int radius = 1, np = 2;  // reaction_diffusion
  if (user_stencil && !user_stencil_info(o->rhs, &radius, &np)) return set_err(ODE_B200_E_INVALID, "unknown stencil %d", o->rhs);
  if (user_field && !user_field_info(o->rhs, &np)) return set_err(ODE_B200_E_INVALID, "unknown field %d", o->rhs);
  MethodInfo mi;
  if (!(field ? pick_field_method(o->method, &mi) : pick_method(o->method, radius, &mi))) {
    if (o->method >= 0 && o->method <= ODE_B200_M_RK4)
      return set_err(ODE_B200_E_NOT_ADAPTIVE, "Can only use this solver with an adaptive RK Butcher table");
    return set_err(ODE_B200_E_UNSUPPORTED, "method %d has no large-system kernel", o->method);
  }
  if (!field && (XS_EDGE + 4 * mi.H > XCHG || mi.VAL < LARGE_THREADS * LARGE_P / 2))
    return set_err(ODE_B200_E_UNSUPPORTED, "stencil radius %d with method %d needs %d ghost points per side: too wide for the tile",
                   radius, o->method, mi.H);
  if (user_stencil) {  // same geometry, kernels compiled at run time for the registered point functor
    ODEB200_CUDA(cudaSetDevice(o->device));
    if (!user_stencil_kernels(o->rhs, o->method, &mi.attempt, &mi.attempt_literal, &mi.hinit_f0, &mi.hinit_f1))
      return ODE_B200_E_INVALID;
  }
  if (user_field) {  // one kernel per stage, compiled at run time for the registered gather functor
    ODEB200_CUDA(cudaSetDevice(o->device));
    const void* k[FIELD_MAX_STAGES + 4] = {};
    if (!user_field_kernels(o->rhs, o->method, mi.S, k)) return ODE_B200_E_INVALID;
    mi.f_first = k[0];
    for (int st = 2; st <= mi.S; st++) mi.f_stage[st] = k[st - 1];
    mi.f_f1 = k[mi.S];
    mi.hinit_f0 = k[mi.S + 1];
    mi.hinit_f1 = k[mi.S + 2];
  }
  if (np > 0 && (!params || o->n_params < np))
    return set_err(ODE_B200_E_INVALID, "the right-hand side needs %d params (reaction_diffusion: {kappa, r})", np);
  if (world < 1 || rank < 0 || rank >= world) return set_err(ODE_B200_E_INVALID, "bad rank/world");
  if (field && world != 1)
    return set_err(ODE_B200_E_UNSUPPORTED, "a general field right-hand side gathers from the whole state: one slab (world = 1) only");
  // (below 32 components the reference's norm is a plain sequential sum, LinearAlgebra.generic_norm2; such
  // systems belong to the ensemble path, which restates it)
  if (n_local < 2 * mi.H || n_global < 32 || n_local > n_global || n_local < 1)
    return set_err(ODE_B200_E_INVALID, "slab too small: n_local %lld (needs >= %d), n_global %lld (needs >= 32)",
                   (long long)n_local, 2 * mi.H, (long long)n_global);
  if (o->points != ODE_B200_POINTS_ALL && o->points != ODE_B200_POINTS_FINAL && o->points != ODE_B200_POINTS_SPECIFIED)
    return set_err(ODE_B200_E_POINTS, "Unrecognized option points==%d", o->points);
  if (!(t0 == t0) || !(tend == tend)) return set_err(ODE_B200_E_INVALID, "NaN in tspan");
  const double tdir = tend > t0 ? 1.0 : (tend < t0 ? -1.0 : 0.0);
  if (tdir == 0.0) return set_err(ODE_B200_E_ZERO_SPAN, "Zero time span");
  if (o->initstep != 0.0 && ((o->initstep > 0 ? 1.0 : -1.0) != tdir))
    return set_err(ODE_B200_E_INITSTEP, "initstep has wrong sign.");
  ODEB200_CUDA(cudaSetDevice(o->device));
  ode_b200_slab* s = new (std::nothrow) ode_b200_slab();
  if (!s) return set_err(ODE_B200_E_NOMEM, "out of memory");
  s->device = o->device;
  s->rank = rank;
  s->world = world;
  s->n = n_local;
  s->n_global = n_global;
  s->mi = mi;
  s->stream = (cudaStream_t)stream;
  s->tiles = (n_local + mi.VAL - 1) / mi.VAL;
  if (!mi.field) {
    s->grid = persistent_grid(mi.attempt, o->device, s->tiles);
    s->grid_literal = persistent_grid(mi.attempt_literal, o->device, s->tiles);
  }
  memset(&s->sc, 0, sizeof(s->sc));
  s->np = np;
  for (int q = 0; q < np; q++) s->sc.p[q] = params[q];
  s->sc.n_global = n_global;
  {  // contiguous slabs in rank order, the first n_global % world ranks one point longer (what shard_bounds of
     // the host mirror does); ode_b200_slab_set_offset overrides
    const long long base = n_global / world, rem = n_global % world;
    s->sc.i0 = rank * base + (rank < rem ? rank : rem);
  }
  // (one tile of zero padding behind the slab: the bulk-copy staging of large.cuh reads whole tile rows)
  const size_t padded = (size_t)(n_local + 2 * mi.H + 8 + LARGE_THREADS * LARGE_P) * sizeof(double);
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, o->device);
  s->init_blocks = sms * 8;
  bool ok = true;
  for (int i = 0; i < 2 && ok; i++) {
    ok = ok && cudaMalloc(&s->Y[i], padded) == cudaSuccess;
    ok = ok && cudaMalloc(&s->K[i], padded) == cudaSuccess;
  }
  if (mi.field) {  // middle stages k_2 .. k_{S-1} and the two stage-argument buffers
    long long fb = (n_local + FIELD_THREADS - 1) / FIELD_THREADS;
    s->field_grid = (int)(fb < (long long)sms * 8 ? (fb < 1 ? 1 : fb) : (long long)sms * 8);
    s->grid = s->field_grid;  // the error partials are written by the last stage kernel
    for (int q = 1; q <= mi.S - 2 && ok; q++) ok = ok && cudaMalloc(&s->KS[q], padded) == cudaSuccess;
    for (int q = 0; q < 2 && ok; q++) ok = ok && cudaMalloc(&s->T[q], padded) == cudaSuccess;
  }
  ok = ok && cudaMalloc(&s->ctl, sizeof(LargeCtl)) == cudaSuccess;
  // per-CTA error partials: the fast and the literal attempt kernels may get different occupancies
  ok = ok && cudaMalloc(&s->part, (size_t)(s->grid > s->grid_literal ? s->grid : s->grid_literal) * 4 * sizeof(double)) == cudaSuccess;
  ok = ok && cudaMalloc(&s->ipart, (size_t)s->init_blocks * 2 * sizeof(double)) == cudaSuccess;
  if (!ok) {
    cudaGetLastError();
    ode_b200_slab_free(s);
    return set_err(ODE_B200_E_NOMEM, "cudaMalloc failed for a slab of %lld points", (long long)n_local);
  }
  for (int i = 0; i < 2; i++) {
    cudaMemsetAsync(s->Y[i], 0, padded, s->stream);
    cudaMemsetAsync(s->K[i], 0, padded, s->stream);
  }
  LargeCtl c;
  memset(&c, 0, sizeof(c));
  c.tdir = tdir;
  c.tstart = t0;
  c.tend = tend;
  c.t = t0;
  c.reltol = o->reltol;
  c.abstol = o->abstol;
  c.minstep = o->minstep;
  c.maxstep = o->maxstep;
  c.initstep = o->initstep;
  c.max_steps = o->max_steps > 0 ? o->max_steps : 100000000LL;
  c.order = mi.order;
  c.t_final = t0;
  s->host_ctl = c;
  ODEB200_CUDA(cudaMemcpyAsync(s->ctl, &s->host_ctl, sizeof(LargeCtl), cudaMemcpyHostToDevice, s->stream));
  ODEB200_CUDA(cudaStreamSynchronize(s->stream));
  *out = s;
  return 0;
}

// (re)load options, parameters and the time span into an existing session
static int slab_configure(ode_b200_slab* s, const ode_b200_opts* o, const double* params, double t0, double tend) {
  if (!(t0 == t0) || !(tend == tend)) return set_err(ODE_B200_E_INVALID, "NaN in tspan");
  const double tdir = tend > t0 ? 1.0 : (tend < t0 ? -1.0 : 0.0);
  if (tdir == 0.0) return set_err(ODE_B200_E_ZERO_SPAN, "Zero time span");
  if (o->initstep != 0.0 && ((o->initstep > 0 ? 1.0 : -1.0) != tdir))
    return set_err(ODE_B200_E_INITSTEP, "initstep has wrong sign.");
  if (!(o->reltol == o->reltol) || !(o->abstol == o->abstol) || !(o->minstep == o->minstep) ||
      !(o->maxstep == o->maxstep) || !(o->initstep == o->initstep))
    return set_err(ODE_B200_E_INVALID, "NaN in reltol/abstol/minstep/maxstep/initstep");
  if (s->np > 0 && (!params || o->n_params < s->np)) return set_err(ODE_B200_E_INVALID, "the stencil RHS needs %d params", s->np);
  for (int q = 0; q < s->np; q++) s->sc.p[q] = params[q];
  LargeCtl c;
  memset(&c, 0, sizeof(c));
  c.tdir = tdir;
  c.tstart = t0;
  c.tend = tend;
  c.t = t0;
  c.reltol = o->reltol;
  c.abstol = o->abstol;
  c.minstep = o->minstep;
  c.maxstep = o->maxstep;
  c.initstep = o->initstep;
  c.max_steps = o->max_steps > 0 ? o->max_steps : 100000000LL;
  c.order = s->mi.order;
  c.t_final = t0;
  c.trace_cap = s->trace ? s->trace_cap : 0;
  s->host_ctl = c;
  ODEB200_CUDA(cudaMemcpyAsync(s->ctl, &s->host_ctl, sizeof(LargeCtl), cudaMemcpyHostToDevice, s->stream));
  ODEB200_CUDA(cudaStreamSynchronize(s->stream));
  return 0;
}

int ode_b200_slab_reset(ode_b200_slab* s, double t0, double tend) {
  int rc = slab_check(s);
  if (rc) return rc;
  if (!(t0 == t0) || !(tend == tend)) return set_err(ODE_B200_E_INVALID, "NaN in tspan");
  const double tdir = tend > t0 ? 1.0 : (tend < t0 ? -1.0 : 0.0);
  if (tdir == 0.0) return set_err(ODE_B200_E_ZERO_SPAN, "Zero time span");
  LargeCtl c = s->host_ctl;
  const double initstep = c.initstep;
  if (initstep != 0.0 && ((initstep > 0 ? 1.0 : -1.0) != tdir)) return set_err(ODE_B200_E_INITSTEP, "initstep has wrong sign.");
  LargeCtl z;
  memset(&z, 0, sizeof(z));
  z.tdir = tdir;
  z.tstart = t0;
  z.tend = tend;
  z.t = t0;
  z.t_final = t0;
  z.reltol = c.reltol;
  z.abstol = c.abstol;
  z.minstep = c.minstep;
  z.maxstep = c.maxstep;
  z.initstep = c.initstep;
  z.max_steps = c.max_steps;
  z.order = c.order;
  z.trace_cap = s->trace ? s->trace_cap : 0;
  s->host_ctl = z;
  ODEB200_CUDA(cudaMemcpyAsync(s->ctl, &s->host_ctl, sizeof(LargeCtl), cudaMemcpyHostToDevice, s->stream));
  ODEB200_CUDA(cudaStreamSynchronize(s->stream));  // host_ctl is reused by poll
  return 0;
}

int ode_b200_slab_set_offset(ode_b200_slab* s, int64_t first_global_index) {
  int rc = slab_check(s);
  if (rc) return rc;
  if (first_global_index < 0 || first_global_index >= s->n_global) return set_err(ODE_B200_E_INVALID, "offset outside the ring");
  s->sc.i0 = first_global_index;
  return 0;
}

int ode_b200_slab_set_fused(ode_b200_slab* s, int32_t on) {
  int rc = slab_check(s);
  if (rc) return rc;
  if (on && ((s->world != 1 && !s->peer_on) || s->mi.field))
    return set_err(ODE_B200_E_UNSUPPORTED, "the step controller can run inside the attempt kernel for a single stencil slab, or "
                                           "for slabs connected over peer memory (ode_b200_slab_peer_connect)");
  s->fused = on != 0;
  return 0;
}

int ode_b200_slab_set_trace(ode_b200_slab* s, double* trace_dev, int64_t trace_cap) {
  int rc = slab_check(s);
  if (rc) return rc;
  s->trace = trace_dev;
  s->trace_cap = trace_cap;
  long long v[2] = {0, trace_cap};
  ODEB200_CUDA(cudaMemcpyAsync(&s->ctl->trace_n, v, sizeof(v), cudaMemcpyHostToDevice, s->stream));
  ODEB200_CUDA(cudaStreamSynchronize(s->stream));
  return 0;
}

int ode_b200_slab_set_state(ode_b200_slab* s, const double* y_local_dev) {
  int rc = slab_check(s);
  if (rc) return rc;
  ODEB200_CUDA(cudaMemcpyAsync(s->Y[0] + s->mi.H, y_local_dev, (size_t)s->n * 8, cudaMemcpyDeviceToDevice, s->stream));
  if (s->out && (!s->out_times || s->out_cap > 0))  // ys[1] = y0, src/runge_kutta.jl:280
    ODEB200_CUDA(cudaMemcpyAsync(s->out, y_local_dev, (size_t)s->n * 8, cudaMemcpyDeviceToDevice, s->stream));
  if (s->out_times && s->out_cap > 0)  // tspan = [tstart] :285
    ODEB200_CUDA(cudaMemcpyAsync(s->out_times, &s->host_ctl.tstart, 8, cudaMemcpyHostToDevice, s->stream));
  return 0;
}

int ode_b200_slab_set_output(ode_b200_slab* s, const double* tspan_host, int32_t n_tspan, double* out_dev) {
  int rc = slab_check(s);
  if (rc) return rc;
  if (s->tq) {
    cudaFree(s->tq);
    s->tq = nullptr;
  }
  s->n_tq = 0;
  s->out = nullptr;
  s->out_times = nullptr;
  s->out_cap = 0;
  if (!tspan_host || n_tspan < 2 || !out_dev) return 0;  // output disabled
  ODEB200_CUDA(cudaMalloc(&s->tq, (size_t)n_tspan * 8));
  ODEB200_CUDA(cudaMemcpyAsync(s->tq, tspan_host, (size_t)n_tspan * 8, cudaMemcpyHostToDevice, s->stream));
  ODEB200_CUDA(cudaStreamSynchronize(s->stream));
  s->n_tq = n_tspan;
  s->out = out_dev;
  return 0;
}

int ode_b200_slab_set_output_all(ode_b200_slab* s, const double* tspan_host, int32_t n_tspan, double* out_dev,
                                 double* out_times_dev, int64_t cap_rows) {
  if (!out_times_dev || cap_rows < 1) return set_err(ODE_B200_E_INVALID, "points=:all output needs a times buffer and cap_rows >= 1");
  int rc = ode_b200_slab_set_output(s, tspan_host, n_tspan, out_dev);
  if (rc) return rc;
  s->out_times = out_times_dev;
  s->out_cap = cap_rows;
  return 0;
}
int ode_b200_slab_get_state(ode_b200_slab* s, double* y_local_dev) {
  int rc = slab_check(s);
  if (rc) return rc;
  int par = 0;
  ODEB200_CUDA(cudaMemcpyAsync(&par, &s->ctl->par, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
  ODEB200_CUDA(cudaStreamSynchronize(s->stream));
  ODEB200_CUDA(cudaMemcpyAsync(y_local_dev, s->Y[par & 1] + s->mi.H, (size_t)s->n * 8, cudaMemcpyDeviceToDevice, s->stream));
  return 0;
}

// phase 0: pack y edges -> xchg_send            [exchange]
// phase 1: y ghosts; f0 = F(t0,y) -> k1; max|y|, max|f0|, f0 edges -> xchg_send   [exchange]
// phase 2: h0; k1 ghosts; f1 = F(t0+tdir*h0, y + tdir*h0*f0); max|f1-f0| -> xchg_send   [exchange]
// phase 3: dt, islaststep
int ode_b200_slab_init_phase(ode_b200_slab* s, int phase, double* xsend, const double* xrecv) {
  int rc = slab_check(s);
  if (rc) return rc;
  const int H = s->mi.H;
  cudaStream_t st = s->stream;
  if (s->mi.field && phase >= 0 && phase <= 2) {  // gather right-hand side: no ghost cells, x1 is materialised
    const double* y = s->Y[0];
    double* f0 = s->K[0];
    const LargeCtl* c = s->ctl;
    if (phase == 1) {
      void* a[] = {(void*)&y, (void*)&f0, (void*)&c, (void*)&s->ipart, (void*)&s->n, (void*)&s->sc};
      ODEB200_CUDA(cudaLaunchKernel(s->mi.hinit_f0, dim3(s->init_blocks), dim3(256), a, 0, st));
      reduce_max_kernel<<<1, 32, 0, st>>>(s->ipart, s->init_blocks, 2, xsend);
      note_launch();
      note_launch();
    } else if (phase == 2) {
      hinit_control_kernel<<<1, 32, 0, st>>>(s->ctl, xrecv, s->world, 0, s->mi.order);
      field_hinit_x1_kernel<<<s->init_blocks, 256, 0, st>>>(y, f0, c, s->T[0], s->n);
      const double* x1 = s->T[0];
      const double* f0c = f0;
      void* a[] = {(void*)&x1, (void*)&f0c, (void*)&c, (void*)&s->ipart, (void*)&s->n, (void*)&s->sc};
      ODEB200_CUDA(cudaLaunchKernel(s->mi.hinit_f1, dim3(s->init_blocks), dim3(256), a, 0, st));
      reduce_max_kernel<<<1, 32, 0, st>>>(s->ipart, s->init_blocks, 1, xsend);
      for (int i = 0; i < 4; i++) note_launch();
    }
    ODEB200_CUDA(cudaGetLastError());
    return 0;
  }
  switch (phase) {
    case 0:
      pack_edges_kernel<<<1, 32, 0, st>>>(s->Y[0], xsend, s->n, H, 0);
      note_launch();
      break;
    case 1:
      large_ghost_kernel<<<1, 32, 0, st>>>(s->Y[0], xrecv, s->n, s->rank, s->world, H, 0);
      {
        const double* y = s->Y[0];
        double* f0 = s->K[0];
        int Hh = H;
        const LargeCtl* c = s->ctl;
        void* a[] = {(void*)&y, (void*)&f0, (void*)&c, (void*)&s->ipart, (void*)&s->n, (void*)&Hh, (void*)&s->sc};
        ODEB200_CUDA(cudaLaunchKernel(s->mi.hinit_f0, dim3(s->init_blocks), dim3(256), a, 0, st));
      }
      reduce_max_kernel<<<1, 32, 0, st>>>(s->ipart, s->init_blocks, 2, xsend);
      pack_edges_kernel<<<1, 32, 0, st>>>(s->K[0], xsend, s->n, H, 1);
      for (int i = 0; i < 4; i++) note_launch();
      break;
    case 2:
      hinit_control_kernel<<<1, 32, 0, st>>>(s->ctl, xrecv, s->world, 0, s->mi.order);
      large_ghost_kernel<<<1, 32, 0, st>>>(s->K[0], xrecv, s->n, s->rank, s->world, H, 1);
      {
        const double* y = s->Y[0];
        const double* f0 = s->K[0];
        const LargeCtl* c = s->ctl;
        int Hh = H;
        void* a[] = {(void*)&y, (void*)&f0, (void*)&c, (void*)&s->ipart, (void*)&s->n, (void*)&Hh, (void*)&s->sc};
        ODEB200_CUDA(cudaLaunchKernel(s->mi.hinit_f1, dim3(s->init_blocks), dim3(256), a, 0, st));
      }
      reduce_max_kernel<<<1, 32, 0, st>>>(s->ipart, s->init_blocks, 1, xsend);
      for (int i = 0; i < 4; i++) note_launch();
      break;
    case 3:
      hinit_control_kernel<<<1, 32, 0, st>>>(s->ctl, xrecv, s->world, 1, s->mi.order);
      note_launch();
      break;
    default:
      return set_err(ODE_B200_E_INVALID, "init phase must be 0..3");
  }
  ODEB200_CUDA(cudaGetLastError());
  return 0;
}

static ControlArgs control_args(const ode_b200_slab* s, const double* xrecv) {
  ControlArgs C;
  C.Y[0] = s->Y[0];
  C.Y[1] = s->Y[1];
  C.K[0] = s->K[0];
  C.K[1] = s->K[1];
  C.ctl = s->ctl;
  C.xrecv = xrecv;
  C.trace = s->trace;
  C.n = s->n;
  C.rank = s->rank;
  C.world = s->world;
  C.H = s->mi.H;
  C.order = s->mi.order;
  C.tq = s->out ? s->tq : nullptr;
  C.n_tq = s->n_tq;
  C.out_times = s->out ? s->out_times : nullptr;
  C.out_cap = s->out_cap;
  return C;
}

int ode_b200_slab_attempt(ode_b200_slab* s, double* xsend) {
  int rc = slab_check(s);
  if (rc) return rc;
  if (s->mi.field) {  // one kernel per stage (field.cuh)
    FieldArgs F;
    memset(&F, 0, sizeof(F));
    for (int q = 0; q < 2; q++) {
      F.Y[q] = s->Y[q];
      F.K[q] = s->K[q];
      F.T[q] = s->T[q];
    }
    for (int q = 0; q < FIELD_MAX_STAGES; q++) F.KS[q] = s->KS[q];
    F.ctl = s->ctl;
    F.part = s->part;
    F.xsend = xsend;
    F.n = s->n;
    F.sc = s->sc;
    void* kargs[] = {(void*)&F};
    const dim3 g((unsigned)s->field_grid), b(FIELD_THREADS);
    ODEB200_CUDA(cudaLaunchKernel(s->mi.f_first, g, b, kargs, 0, s->stream));
    note_launch();
    for (int st = 2; st <= s->mi.S; st++) {
      ODEB200_CUDA(cudaLaunchKernel(s->mi.f_stage[st], g, b, kargs, 0, s->stream));
      note_launch();
    }
    if (!s->mi.fsal) {
      ODEB200_CUDA(cudaLaunchKernel(s->mi.f_f1, g, b, kargs, 0, s->stream));
      note_launch();
    }
    return 0;
  }
  LargeArgs A;
  A.Y[0] = s->Y[0];
  A.Y[1] = s->Y[1];
  A.K[0] = s->K[0];
  A.K[1] = s->K[1];
  A.ctl = s->ctl;
  A.part = s->part;
  A.xsend = xsend;
  A.n = s->n;
  A.tiles = s->tiles;
  A.sc = s->sc;
  A.fused = s->fused ? 1 : 0;
  A.ctl_args = control_args(s, xsend);  // fused mode: the partials are read where the last CTA wrote them
  memset(&A.peer, 0, sizeof(A.peer));
  if (s->fused && s->peer_on) A.peer = s->px;  // world > 1: the last CTA all-gathers the records over peer memory
  void* kargs[] = {(void*)&A};
  ODEB200_CUDA(cudaLaunchKernel(s->mi.attempt, dim3((unsigned)s->grid), dim3(LARGE_THREADS), kargs, 0, s->stream));
  // the literal repeat of an attempt that went non-finite (ctl->redo); returns at once otherwise
  ODEB200_CUDA(cudaLaunchKernel(s->mi.attempt_literal, dim3((unsigned)s->grid_literal), dim3(LARGE_THREADS), kargs, 0, s->stream));
  note_launch();
  note_launch();
  return 0;
}

int ode_b200_slab_control(ode_b200_slab* s, const double* xrecv) {
  int rc = slab_check(s);
  if (rc) return rc;
  if (!s->fused) {
    const ControlArgs C = control_args(s, xrecv);
    large_control_kernel<<<1, 32, 0, s->stream>>>(C);
    note_launch();
  }
  if (s->out) {  // before the next attempt overwrites the step-start arrays
    large_hermite_kernel<<<s->init_blocks, 256, 0, s->stream>>>(s->Y[0], s->Y[1], s->K[0], s->K[1], s->ctl, s->tq,
                                                                 s->out, s->n, s->mi.H, s->out_cap,
                                                                 s->out_times ? 1 : 0);
    note_launch();
  }
  ODEB200_CUDA(cudaGetLastError());
  return 0;
}

// ctl_out[16]: {done, status, n_accept, n_reject, t, dt, err, t_final, n_rhs, attempts, trace_n, ghost H, out_rows, -, -, -}
int ode_b200_slab_poll(ode_b200_slab* s, double* ctl_out) {
  int rc = slab_check(s);
  if (rc) return rc;
  ODEB200_CUDA(cudaMemcpyAsync(&s->host_ctl, s->ctl, sizeof(LargeCtl), cudaMemcpyDeviceToHost, s->stream));
  ODEB200_CUDA(cudaStreamSynchronize(s->stream));
  const LargeCtl& c = s->host_ctl;
  if (ctl_out) {
    ctl_out[0] = c.done;
    ctl_out[1] = c.status;
    ctl_out[2] = (double)c.nacc;
    ctl_out[3] = (double)c.nrej;
    ctl_out[4] = c.t;
    ctl_out[5] = c.dt;
    ctl_out[6] = c.err;
    ctl_out[7] = c.t_final;
    // RHS sweeps: 2 in hinit, S-1 per attempt, +1 per accepted step when the tableau is not FSAL
    ctl_out[8] = (double)(c.nrhs + c.attempts * (s->mi.S - 1) + (s->mi.fsal ? 0 : c.nacc));
    ctl_out[9] = (double)c.attempts;
    ctl_out[10] = (double)c.trace_n;
    ctl_out[11] = (double)s->mi.H;
    ctl_out[12] = (double)c.out_rows;
  }
  return 0;
}

void ode_b200_slab_free(ode_b200_slab* s) {
  if (!s) return;
  cudaSetDevice(s->device);
  for (int i = 0; i < 2; i++) {
    if (s->Y[i]) cudaFree(s->Y[i]);
    if (s->K[i]) cudaFree(s->K[i]);
  }
  for (int q = 0; q < FIELD_MAX_STAGES; q++)
    if (s->KS[q]) cudaFree(s->KS[q]);
  for (int q = 0; q < 2; q++)
    if (s->T[q]) cudaFree(s->T[q]);
  if (s->ctl) cudaFree(s->ctl);
  if (s->part) cudaFree(s->part);
  if (s->ipart) cudaFree(s->ipart);
  if (s->tq) cudaFree(s->tq);
  if (s->xs) cudaFree(s->xs);
  if (s->xr) cudaFree(s->xr);
  if (s->pin_ctl) cudaFreeHost(s->pin_ctl);
  for (int q = 0; q < 2; q++)
    if (s->ev_poll[q]) cudaEventDestroy(s->ev_poll[q]);
  for (int r = 0; r < PEER_MAX; r++)
    if (s->peer_ipc[r]) cudaIpcCloseMemHandle(s->peer_ipc[r]);
  if (s->peer_block) cudaFree(s->peer_block);
  delete s;
}

// ---- exchange over peer memory ---------------------------------------------------------------------------------
static size_t peer_block_bytes(int world) { return ((size_t)2 * world * XCHG + 2 * world + 8) * sizeof(double); }

int ode_b200_slab_peer_alloc(ode_b200_slab* s, void* ipc_handle_out, void** local_ptr_out) {
  int rc = slab_check(s);
  if (rc) return rc;
  if (s->world > PEER_MAX) return set_err(ODE_B200_E_UNSUPPORTED, "at most %d slabs can exchange over peer memory", PEER_MAX);
  if (s->mi.field) return set_err(ODE_B200_E_UNSUPPORTED, "a general field right-hand side runs on one slab");
  if (!s->peer_block) {
    ODEB200_CUDA(cudaMalloc(&s->peer_block, peer_block_bytes(s->world)));
    ODEB200_CUDA(cudaMemset(s->peer_block, 0, peer_block_bytes(s->world)));  // flags and exchange counter start at 0
    ODEB200_CUDA(cudaDeviceSynchronize());
    s->peer_epoch = reinterpret_cast<unsigned long long*>(s->peer_block + 2LL * s->world * XCHG) + 2 * s->world;
  }
  if (ipc_handle_out) {
    static_assert(sizeof(cudaIpcMemHandle_t) == ODE_B200_PEER_HANDLE_BYTES, "handle size");
    cudaIpcMemHandle_t h;
    ODEB200_CUDA(cudaIpcGetMemHandle(&h, s->peer_block));
    memcpy(ipc_handle_out, &h, sizeof(h));
  }
  if (local_ptr_out) *local_ptr_out = s->peer_block;
  return 0;
}

int ode_b200_slab_peer_connect(ode_b200_slab* s, const void* ipc_handles, void* const* ptrs) {
  int rc = slab_check(s);
  if (rc) return rc;
  if (!s->peer_block) return set_err(ODE_B200_E_INVALID, "ode_b200_slab_peer_alloc comes first");
  if (!ipc_handles && !ptrs) return set_err(ODE_B200_E_INVALID, "peer handles or peer pointers are required");
  memset(&s->px, 0, sizeof(s->px));
  for (int r = 0; r < s->world; r++) {
    if (r == s->rank) {
      s->px.block[r] = s->peer_block;
    } else if (ptrs) {  // same process: the caller enabled peer access between the devices
      if (!ptrs[r]) return set_err(ODE_B200_E_INVALID, "peer pointer %d is null", r);
      s->px.block[r] = (double*)ptrs[r];
    } else {
      if (s->peer_ipc[r]) {
        cudaIpcCloseMemHandle(s->peer_ipc[r]);
        s->peer_ipc[r] = nullptr;
      }
      cudaIpcMemHandle_t h;
      memcpy(&h, (const char*)ipc_handles + (size_t)r * ODE_B200_PEER_HANDLE_BYTES, sizeof(h));
      void* p = nullptr;
      ODEB200_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
      s->peer_ipc[r] = p;
      s->px.block[r] = (double*)p;
    }
  }
  s->px.mine = s->peer_block;
  s->px.epoch = s->peer_epoch;
  s->px.world = s->world;
  s->px.rank = s->rank;
  long long ms = 5000;
  if (const char* e = getenv("ODE_B200_PEER_TIMEOUT_MS")) ms = atoll(e) > 0 ? atoll(e) : ms;
  s->px.timeout_ns = ms * 1000000LL;
  s->peer_on = true;
  s->fused = true;
  if (const char* e = getenv("ODE_B200_PEER_FUSED")) s->fused = atoi(e) != 0;
  return 0;
}

// recv <- all_gather(send) in stream order, for the exchanges the session can do itself
static int slab_exchange(ode_b200_slab* s, double* xsend, double* xrecv) {
  if (s->world == 1) {
    ODEB200_CUDA(cudaMemcpyAsync(xrecv, xsend, XCHG * 8, cudaMemcpyDeviceToDevice, s->stream));
    return 0;
  }
  if (!s->peer_on)
    return set_err(ODE_B200_E_INVALID, "slabs of a decomposed grid that are not connected over peer memory are driven step by step "
                                       "(ode_b200_slab_attempt / _control) with the caller's all-gather");
  PeerXchg P = s->px;
  P.gathered = xrecv;
  peer_exchange_kernel<<<1, 128, 0, s->stream>>>(P, xsend, XS_EDGE + 4 * s->mi.H, s->ctl);
  note_launch();
  ODEB200_CUDA(cudaGetLastError());
  return 0;
}

// The whole integration loop of `count` sessions of THIS process, in lockstep: hinit phases, then batches of 4
// attempts (captured once per session into a CUDA graph and replayed; the done flag of batch k is looked at while
// batch k+1 is already queued, and attempts issued after the end are no-ops).  count = 1: one slab (world = 1), or this process's slab of a grid
// decomposed across processes that are connected over peer memory.  count > 1: all slabs live in this process, one
// per GPU; every launch is asynchronous, so one host thread keeps all devices busy, and nothing here blocks on a
// device before the matching work of every other device has been enqueued.
static int slabs_run(ode_b200_slab* const* ss, int count, double (*polls)[16], float* ms_out) {
  int rc = 0;
  for (int g = 0; g < count && !rc; g++) {
    ode_b200_slab* s = ss[g];
    rc = slab_check(s);
    if (!rc && !s->xs) {
      ODEB200_CUDA(cudaMalloc(&s->xs, XCHG * 8));
      ODEB200_CUDA(cudaMalloc(&s->xr, (size_t)s->world * XCHG * 8));
    }
    if (!rc && s->world > 1 && !s->peer_on) rc = slab_exchange(s, nullptr, nullptr);  // (sets the error)
  }
  if (rc) return rc;
  std::vector<cudaEvent_t> e0(count, nullptr), e1(count, nullptr);
  std::vector<cudaGraphExec_t> gexec(count, nullptr);
  std::vector<long long> batch_launches(count, 0);
  auto cleanup = [&]() {
    for (int g = 0; g < count; g++) {
      cudaSetDevice(ss[g]->device);
      if (gexec[g]) cudaGraphExecDestroy(gexec[g]);
      if (e0[g]) cudaEventDestroy(e0[g]);
      if (e1[g]) cudaEventDestroy(e1[g]);
    }
  };
#define RUN_TRY(x)      \
  do {                  \
    if (!rc) rc = (x);  \
  } while (0)
#define RUN_CUDA(x)                                                                                  \
  do {                                                                                               \
    if (!rc) {                                                                                       \
      cudaError_t e__ = (x);                                                                         \
      if (e__ != cudaSuccess) rc = set_err(ODE_B200_E_CUDA, "%s: %s", #x, cudaGetErrorString(e__));  \
    }                                                                                                \
  } while (0)
  for (int g = 0; g < count && !rc; g++) {
    ode_b200_slab* s = ss[g];
    RUN_CUDA(cudaSetDevice(s->device));
    RUN_CUDA(cudaMemsetAsync(s->xs, 0, XCHG * 8, s->stream));
    RUN_CUDA(cudaEventCreate(&e0[g]));
    RUN_CUDA(cudaEventCreate(&e1[g]));
    RUN_CUDA(cudaEventRecord(e0[g], s->stream));
  }
  for (int ph = 0; ph < 4 && !rc; ph++) {
    for (int g = 0; g < count && !rc; g++) {
      ode_b200_slab* s = ss[g];
      RUN_CUDA(cudaSetDevice(s->device));
      RUN_TRY(ode_b200_slab_init_phase(s, ph, s->xs, s->xr));
      if (ph < 3) RUN_TRY(slab_exchange(s, s->xs, s->xr));
    }
  }
  const int batch = 4;  // (batches of 16 were slower: the graph costs more to instantiate than the fewer polls save)
  auto enqueue_batch = [&](ode_b200_slab* s) {
    for (int q = 0; q < batch && !rc; q++) {
      RUN_TRY(ode_b200_slab_attempt(s, s->xs));
      if (!s->fused) RUN_TRY(slab_exchange(s, s->xs, s->xr));
      RUN_TRY(ode_b200_slab_control(s, s->xr));
    }
  };
  // ODE_B200_NO_GRAPH=1 (or a failed capture) issues the launches directly
  if (!rc && !getenv("ODE_B200_NO_GRAPH")) {
    for (int g = 0; g < count && !rc; g++) {
      ode_b200_slab* s = ss[g];
      RUN_CUDA(cudaSetDevice(s->device));
      cudaGraph_t graph = nullptr;
      const long long before = ode_b200_launch_count(0);
      if (cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
        enqueue_batch(s);
        const cudaError_t ce = cudaStreamEndCapture(s->stream, &graph);
        if (ce != cudaSuccess || rc || !graph || cudaGraphInstantiate(&gexec[g], graph, 0) != cudaSuccess) gexec[g] = nullptr;
        if (graph) cudaGraphDestroy(graph);
        if (!gexec[g]) {  // fall back to direct launches
          cudaGetLastError();
          rc = 0;
        }
      } else {
        cudaGetLastError();
      }
      batch_launches[g] = ode_b200_launch_count(0) - before;  // counted once at capture; added per replay below
    }
  }
  // Pipelined polling: batch k+1 is enqueued BEFORE the host looks at the control block batch k left behind (copied
  // into page-locked memory behind batch k, with an event), so the device never waits for the host's round trip --
  // 5-7 us per attempt that matter when an attempt is short (small systems, strong scaling), and, across processes,
  // the ranks' host jitter no longer shows up as a wait for the slowest rank's record.  At most one batch of no-ops
  // is enqueued past the end.
  for (int g = 0; g < count && !rc; g++) {
    ode_b200_slab* s = ss[g];
    RUN_CUDA(cudaSetDevice(s->device));
    if (!s->pin_ctl) {
      RUN_CUDA(cudaHostAlloc((void**)&s->pin_ctl, 2 * sizeof(LargeCtl), cudaHostAllocDefault));
      for (int q = 0; q < 2; q++) RUN_CUDA(cudaEventCreateWithFlags(&s->ev_poll[q], cudaEventDisableTiming));
    }
  }
  bool first_batch = true;
  for (long long k = 0; !rc; k++) {
    for (int g = 0; g < count && !rc; g++) {
      ode_b200_slab* s = ss[g];
      RUN_CUDA(cudaSetDevice(s->device));
      if (gexec[g]) {
        RUN_CUDA(cudaGraphLaunch(gexec[g], s->stream));
        if (!first_batch)
          for (long long q = 0; q < batch_launches[g]; q++) note_launch();
      } else {
        enqueue_batch(s);
      }
      RUN_CUDA(cudaMemcpyAsync(&s->pin_ctl[k & 1], s->ctl, sizeof(LargeCtl), cudaMemcpyDeviceToHost, s->stream));
      RUN_CUDA(cudaEventRecord(s->ev_poll[k & 1], s->stream));
    }
    first_batch = false;
    if (k == 0) continue;  // keep one batch in flight ahead of the poll
    bool all_done = true;
    for (int g = 0; g < count && !rc; g++) {
      ode_b200_slab* s = ss[g];
      RUN_CUDA(cudaSetDevice(s->device));
      RUN_CUDA(cudaEventSynchronize(s->ev_poll[(k - 1) & 1]));
      all_done = all_done && s->pin_ctl[(k - 1) & 1].done != 0;
    }
    if (all_done) break;
  }
  for (int g = 0; g < count && !rc; g++) RUN_TRY(ode_b200_slab_poll(ss[g], polls[g]));  // (waits for the batch in flight)
  float ms = 0.f;
  for (int g = 0; g < count && !rc; g++) {
    ode_b200_slab* s = ss[g];
    RUN_CUDA(cudaSetDevice(s->device));
    RUN_CUDA(cudaEventRecord(e1[g], s->stream));
    RUN_CUDA(cudaEventSynchronize(e1[g]));
    float m = 0.f;
    if (!rc) cudaEventElapsedTime(&m, e0[g], e1[g]);
    ms = m > ms ? m : ms;
  }
  if (ms_out) *ms_out = ms;
  cleanup();
  return rc;
#undef RUN_TRY
#undef RUN_CUDA
}

int ode_b200_slab_run(ode_b200_slab* s, double* ctl_out, double* kernel_ms) {
  int rc = slab_check(s);
  if (rc) return rc;
  double poll[1][16] = {{0}};
  float ms = 0.f;
  if (s->world == 1 && !s->mi.field) s->fused = true;
  rc = slabs_run(&s, 1, poll, &ms);
  if (ctl_out) memcpy(ctl_out, poll[0], sizeof(poll[0]));
  if (kernel_ms) *kernel_ms = ms;
  return rc;
}

static int solve_large_impl(const ode_b200_opts* o, int64_t n, const double* y0, const double* params,
                            const double* tspan, int32_t n_tspan, double* y_out, double* pts_out,
                            ode_b200_large_stats* stats, double* trace, int64_t trace_cap, int64_t* trace_n,
                            double* all_t = nullptr, int64_t all_cap = 0, int64_t* all_rows = nullptr);

// oderk_fixed (src/runge_kutta.jl:176-212) on a dof = N state: ode1 / ode2_midpoint / ode2_heun / ode4 with a
// stencil or field right-hand side, one kernel per stage per step on the tspan grid.
int ode_b200_solve_large_fixed(const ode_b200_opts* o, int64_t n, const double* y0, const double* params,
                               const double* tspan, int32_t n_tspan, double* pts_y, double* y_final,
                               ode_b200_large_stats* stats) {
  if (!o || !y0 || !tspan) return set_err(ODE_B200_E_INVALID, "null argument");
  if (n_tspan < 2) return set_err(ODE_B200_E_INVALID, "tspan needs at least 2 entries");
  int cnt = 0;
  if (cudaGetDeviceCount(&cnt) != cudaSuccess || cnt <= 0) {
    cudaGetLastError();
    return set_err(ODE_B200_E_NO_DEVICE, "no CUDA device available; libode_b200 has no CPU fallback");
  }
  if (o->device < 0 || o->device >= cnt) return set_err(ODE_B200_E_INVALID, "device %d out of range", o->device);
  const int msord = ms_order(o->method);  // ode_ms (Adams-Bashforth, order 1..5): one kernel per step, a ring of derivative arrays
  const bool ms4 = msord > 0;
  const int S = ms4 ? 1 : fixed_stage_count(o->method);
  if (S == 0)
    return set_err(ODE_B200_E_UNSUPPORTED, "method %d is not a fixed-step explicit method with a large-system kernel "
                                           "(the Rosenbrock solvers need a dense N x N Jacobian)", o->method);
  if (n < 32) return set_err(ODE_B200_E_INVALID, "large-system path: n %lld (needs >= 32; smaller systems take the ensemble path)", (long long)n);
  ODEB200_CUDA(cudaSetDevice(o->device));
  const void* k[FIELD_MAX_STAGES] = {};
  int np = 2;
  if (o->rhs == ODE_B200_RHS_REACTION_DIFFUSION || o->rhs == ODE_B200_RHS_RD_FIELD) {
    if (ms4)
      k[0] = (const void*)field_ms_kernel<FieldReactionDiffusion>;
    else
      builtin_fixed_kernels(o->method, k);
  } else if (o->rhs >= ODE_B200_FIELD_USER_BASE) {
    if (!user_field_info(o->rhs, &np)) return set_err(ODE_B200_E_INVALID, "unknown field %d", o->rhs);
    if (!user_fixed_kernels(o->rhs, o->method, S, k)) return ODE_B200_E_INVALID;
  } else if (o->rhs >= ODE_B200_STENCIL_USER_BASE) {
    int radius = 1;
    if (!user_stencil_info(o->rhs, &radius, &np)) return set_err(ODE_B200_E_INVALID, "unknown stencil %d", o->rhs);
    if (!user_fixed_kernels(o->rhs, o->method, S, k)) return ODE_B200_E_INVALID;
  } else {
    return set_err(ODE_B200_E_UNSUPPORTED, "the large-system path takes a stencil or field right-hand side");
  }
  if (np > 0 && (!params || o->n_params < np)) return set_err(ODE_B200_E_INVALID, "the right-hand side needs %d params", np);
  // device buffers: two states, S-1 stages, two stage arguments
  std::vector<double*> bufs;
  auto release = [&]() {
    for (double* b : bufs) cudaFree(b);
  };
  auto alloc = [&](double** p) {
    if (cudaMalloc(p, (size_t)n * sizeof(double)) != cudaSuccess) return false;
    bufs.push_back(*p);
    return true;
  };
  FixedArgs A;
  memset(&A, 0, sizeof(A));
  double *ya = nullptr, *yb = nullptr;
  constexpr int MO = ODE_B200_MS_MAX_ORDER;
  double* hist[MO] = {};  // ode_ms: ring of the last MO derivatives
  bool ok = alloc(&ya) && alloc(&yb);
  if (ms4) {
    for (int q = 0; q < MO && ok; q++) ok = alloc(&hist[q]);
  } else {
    ok = ok && alloc(&A.T[0]) && alloc(&A.T[1]);
    for (int q = 0; q < S - 1 && ok; q++) ok = alloc(&A.KS[q]);
  }
  if (!ok) {
    cudaGetLastError();
    release();
    return set_err(ODE_B200_E_NOMEM, "cudaMalloc failed for a fixed-step system of %lld components", (long long)n);
  }
  cudaStream_t st = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  A.n = n;
  for (int q = 0; q < np; q++) A.sc.p[q] = params[q];
  A.sc.n_global = n;
  A.sc.i0 = 0;
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, o->device);
  long long fb = (n + FIELD_THREADS - 1) / FIELD_THREADS;
  const unsigned grid = (unsigned)(fb < (long long)sms * 8 ? fb : (long long)sms * 8);
  cudaError_t ce = cudaMemcpyAsync(ya, y0, (size_t)n * 8, cudaMemcpyHostToDevice, st);
  if (pts_y) memcpy(pts_y, y0, (size_t)n * 8);  // ys[1] = deepcopy(y0) :193
  cudaEventRecord(e0, st);
  for (int i = 0; i + 1 < n_tspan && ce == cudaSuccess; i++) {  // :201
    A.y = ya;
    A.ynew = yb;
    A.t = tspan[i];
    A.dt = tspan[i + 1] - tspan[i];  // :202
    if (ms4) {  // xdot[i] goes to ring slot i % MO; the older ones are slots i-4 .. i-1
      MsArgs M;
      memset(&M, 0, sizeof(M));
      M.y = ya;
      M.ynew = yb;
      for (int q = 0; q < MO - 1; q++) M.hold[q] = hist[(i + 1 + q) % MO];  // xdot[i-4+q]
      M.hnew = hist[i % MO];
      M.n = n;
      M.t = A.t;
      M.h = A.dt;
      M.so = (i + 1 < msord) ? i + 1 : msord;  // steporder = min(i, order), src/ODE.jl:200
      for (int q = 0; q < MO; q++) {           // h[i]*b[steporder, j] (:205), j = q - (MO - so), oldest derivative first
        const int j = q - (MO - M.so);
        M.hb[q] = j >= 0 ? A.dt * ms_coefficient(M.so, j) : 0.0;
      }
      M.sc = A.sc;
      void* margs[] = {(void*)&M};
      ce = cudaLaunchKernel(k[0], dim3(grid), dim3(FIELD_THREADS), margs, 0, st);
      note_launch();
    } else {
      void* kargs[] = {(void*)&A};
      for (int s = 1; s <= S && ce == cudaSuccess; s++) {
        ce = cudaLaunchKernel(k[s - 1], dim3(grid), dim3(FIELD_THREADS), kargs, 0, st);
        note_launch();
      }
    }
    if (pts_y && ce == cudaSuccess)
      ce = cudaMemcpyAsync(pts_y + (size_t)(i + 1) * n, yb, (size_t)n * 8, cudaMemcpyDeviceToHost, st);
    double* tmp = ya;
    ya = yb;
    yb = tmp;
  }
  cudaEventRecord(e1, st);
  if (y_final && ce == cudaSuccess) ce = cudaMemcpyAsync(y_final, ya, (size_t)n * 8, cudaMemcpyDeviceToHost, st);
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
  float ms = 0.f;
  if (ce == cudaSuccess) cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaStreamDestroy(st);
  release();
  if (ce != cudaSuccess) {
    cudaGetLastError();
    return set_err(ODE_B200_E_CUDA, "fixed-step large-system solve failed: %s", cudaGetErrorString(ce));
  }
  if (stats) {
    memset(stats, 0, sizeof(*stats));
    stats->t_final = tspan[n_tspan - 1];
    stats->n_accept = n_tspan - 1;
    stats->n_rhs = (int64_t)(n_tspan - 1) * S;
    stats->kernel_ms = ms;
  }
  return 0;
}

void ode_b200_trim(void) {
  large_cache_release();
  arena_release();
}

int ode_b200_solve_large(const ode_b200_opts* o, int64_t n, const double* y0, const double* params, double t0,
                         double tend, double* y_out, ode_b200_large_stats* stats, double* trace, int64_t trace_cap,
                         int64_t* trace_n) {
  if (o && o->n_gpus > 1)  // fan-out inside the library: devices o->device .. o->device + n_gpus - 1
    return ode_b200_solve_large_multi(o, o->n_gpus, nullptr, n, y0, params, t0, tend, y_out, stats, trace, trace_cap, trace_n);
  const double ts[2] = {t0, tend};
  return solve_large_impl(o, n, y0, params, ts, 2, y_out, nullptr, stats, trace, trace_cap, trace_n);
}

int ode_b200_solve_large_points(const ode_b200_opts* o, int64_t n, const double* y0, const double* params,
                                const double* tspan, int32_t n_tspan, double* pts_y, ode_b200_large_stats* stats,
                                double* trace, int64_t trace_cap, int64_t* trace_n) {
  if (!tspan || n_tspan < 2 || !pts_y) return set_err(ODE_B200_E_INVALID, "tspan (>= 2 entries) and pts_y are required");
  return solve_large_impl(o, n, y0, params, tspan, n_tspan, nullptr, pts_y, stats, trace, trace_cap, trace_n);
}

int ode_b200_solve_large_all(const ode_b200_opts* o, int64_t n, const double* y0, const double* params,
                             const double* tspan, int32_t n_tspan, int64_t cap_rows, double* pts_t, double* pts_y,
                             int64_t* n_rows, ode_b200_large_stats* stats, double* trace, int64_t trace_cap,
                             int64_t* trace_n) {
  if (!tspan || n_tspan < 2 || !pts_y || !pts_t || cap_rows < 1 || !n_rows)
    return set_err(ODE_B200_E_INVALID, "tspan, pts_t, pts_y, n_rows and cap_rows >= 1 are required");
  return solve_large_impl(o, n, y0, params, tspan, n_tspan, nullptr, pts_y, stats, trace, trace_cap, trace_n, pts_t,
                          cap_rows, n_rows);
}

// ODE_B200_GUARD_CHECK=1 (tests; compute-sanitizer is not available on every pool): every state / stage array of a
// session is allocated with a zeroed tail behind the last ghost cell; no kernel may ever write there.  Counts the
// doubles of those tails that are no longer +0.
__global__ void guard_count_kernel(const double* a, long long lo, long long hi, unsigned int* count) {
  for (long long i = lo + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < hi; i += (long long)gridDim.x * blockDim.x)
    if (__double_as_longlong(a[i]) != 0LL) atomicAdd(count, 1u);
}
static int slab_guard_violations(ode_b200_slab* s, long long* out) {
  *out = 0;
  unsigned int* d = nullptr;
  ODEB200_CUDA(cudaMalloc(&d, sizeof(unsigned int)));
  ODEB200_CUDA(cudaMemsetAsync(d, 0, sizeof(unsigned int), s->stream));
  const long long lo = s->n + 2 * s->mi.H, hi = lo + 8 + LARGE_THREADS * LARGE_P;
  const double* arrays[4 + FIELD_MAX_STAGES + 2] = {s->Y[0], s->Y[1], s->K[0], s->K[1]};
  int na = 4;
  for (int q = 0; q < FIELD_MAX_STAGES; q++)
    if (s->KS[q]) arrays[na++] = s->KS[q];
  for (int q = 0; q < 2; q++)
    if (s->T[q]) arrays[na++] = s->T[q];
  for (int q = 0; q < na; q++)
    if (arrays[q]) guard_count_kernel<<<2, 256, 0, s->stream>>>(arrays[q], lo, hi, d);
  unsigned int h = 0;
  ODEB200_CUDA(cudaMemcpyAsync(&h, d, sizeof(h), cudaMemcpyDeviceToHost, s->stream));
  ODEB200_CUDA(cudaStreamSynchronize(s->stream));
  cudaFree(d);
  *out = h;
  return 0;
}

static int solve_large_impl(const ode_b200_opts* o, int64_t n, const double* y0, const double* params,
                            const double* tspan, int32_t n_tspan, double* y_out, double* pts_out,
                            ode_b200_large_stats* stats, double* trace, int64_t trace_cap, int64_t* trace_n,
                            double* all_t, int64_t all_cap, int64_t* all_rows) {
  if (!o || !y0) return set_err(ODE_B200_E_INVALID, "null argument");
  const double t0 = tspan[0], tend = tspan[n_tspan - 1];
  int cnt = 0;
  if (cudaGetDeviceCount(&cnt) != cudaSuccess || cnt <= 0) {
    cudaGetLastError();
    return set_err(ODE_B200_E_NO_DEVICE, "no CUDA device available; libode_b200 has no CPU fallback");
  }
  if (o->device < 0 || o->device >= cnt) return set_err(ODE_B200_E_INVALID, "device %d out of range", o->device);
  ODEB200_CUDA(cudaSetDevice(o->device));
  // Device buffers of the last call are kept per thread and reused when (device, method, n) repeat,
  // so repeated solves (parameter sweeps, the benchmark's e2e arm) do not pay cudaMalloc/cudaFree of
  // 5 arrays of n doubles each time.  ode_b200_trim() releases them.
  LargeCache& cc = g_large_cache;
  int rc = 0;
  if (cc.s && (cc.device != o->device || cc.method != o->method || cc.n != n || cc.rhs != o->rhs)) large_cache_release();
  if (!cc.s) {
    ODEB200_CUDA(cudaStreamCreateWithFlags(&cc.st, cudaStreamNonBlocking));
    rc = ode_b200_slab_create(o, n, n, 0, 1, params, t0, tend, cc.st, &cc.s);
    if (!rc && cudaMalloc(&cc.dy, (size_t)n * 8) != cudaSuccess) rc = set_err(ODE_B200_E_NOMEM, "cudaMalloc failed");
    if (rc) {
      cudaGetLastError();
      large_cache_release();
      return rc;
    }
    cc.device = o->device;
    cc.method = o->method;
    cc.rhs = o->rhs;
    cc.n = n;
  } else {
    if (o->points != ODE_B200_POINTS_ALL && o->points != ODE_B200_POINTS_FINAL && o->points != ODE_B200_POINTS_SPECIFIED)
      return set_err(ODE_B200_E_POINTS, "Unrecognized option points==%d", o->points);
    cc.s->trace = nullptr;
    cc.s->trace_cap = 0;
    rc = slab_configure(cc.s, o, params, t0, tend);
    if (rc) return rc;
  }
  ode_b200_slab* s = cc.s;
  cudaStream_t st = cc.st;
  double *dy = cc.dy, *dtrace = nullptr, *dpts = nullptr;
  double poll[16] = {0};
#define LARGE_TRY(x)  \
  do {                \
    if (!rc) rc = (x); \
  } while (0)
#define LARGE_CUDA(x)                                                                          \
  do {                                                                                         \
    if (!rc) {                                                                                 \
      cudaError_t e__ = (x);                                                                   \
      if (e__ != cudaSuccess) rc = set_err(ODE_B200_E_CUDA, "%s: %s", #x, cudaGetErrorString(e__)); \
    }                                                                                          \
  } while (0)
  LARGE_TRY(ode_b200_slab_set_output(s, nullptr, 0, nullptr));
  if (trace && trace_cap > 0) {
    LARGE_CUDA(cudaMalloc(&dtrace, (size_t)trace_cap * 32));
    LARGE_TRY(ode_b200_slab_set_trace(s, dtrace, trace_cap));
  }
  double* dtimes = nullptr;
  if (pts_out && all_t) {  // points = :all, capped at all_cap rows
    LARGE_CUDA(cudaMalloc(&dpts, (size_t)all_cap * n * 8));
    LARGE_CUDA(cudaMalloc(&dtimes, (size_t)all_cap * 8));
    LARGE_TRY(ode_b200_slab_set_output_all(s, tspan, n_tspan, dpts, dtimes, all_cap));
  } else if (pts_out) {
    LARGE_CUDA(cudaMalloc(&dpts, (size_t)n_tspan * n * 8));
    LARGE_TRY(ode_b200_slab_set_output(s, tspan, n_tspan, dpts));
  }
  LARGE_CUDA(cudaMemcpyAsync(dy, y0, (size_t)n * 8, cudaMemcpyHostToDevice, st));
  LARGE_TRY(ode_b200_slab_set_state(s, dy));
  s->fused = !s->mi.field;  // one slab: accept/reject runs inside the attempt kernel (no exchange, no control launch)
  float run_ms = 0.f;
  double polls[1][16] = {{0}};
  LARGE_TRY(slabs_run(&s, 1, polls, &run_ms));  // hinit phases + graph-replayed attempt batches until done
  memcpy(poll, polls[0], sizeof(poll));
  LARGE_TRY(ode_b200_slab_get_state(s, dy));
  if (y_out) LARGE_CUDA(cudaMemcpyAsync(y_out, dy, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
  if (pts_out && dpts && all_t) {
    const long long rows = (long long)poll[12];
    const long long m = rows < all_cap ? rows : all_cap;
    LARGE_CUDA(cudaMemcpyAsync(pts_out, dpts, (size_t)m * n * 8, cudaMemcpyDeviceToHost, st));
    LARGE_CUDA(cudaMemcpyAsync(all_t, dtimes, (size_t)m * 8, cudaMemcpyDeviceToHost, st));
    if (all_rows) *all_rows = rows;
  } else if (pts_out && dpts) {
    LARGE_CUDA(cudaMemcpyAsync(pts_out, dpts, (size_t)n_tspan * n * 8, cudaMemcpyDeviceToHost, st));
  }
  if (trace && dtrace) {
    long long tn = (long long)poll[10];
    long long m = tn < trace_cap ? tn : trace_cap;
    LARGE_CUDA(cudaMemcpyAsync(trace, dtrace, (size_t)m * 32, cudaMemcpyDeviceToHost, st));
    if (trace_n) *trace_n = tn;
  }
  LARGE_CUDA(cudaStreamSynchronize(st));
  if (!rc && s && getenv("ODE_B200_GUARD_CHECK")) {
    long long bad = 0;
    if (getenv("ODE_B200_GUARD_SELFTEST"))  // (the check checked: poke one double of a tail, the solve must fail)
      LARGE_CUDA(cudaMemsetAsync(s->K[1] + s->n + 2 * s->mi.H + 3, 0x3f, 8, st));
    LARGE_TRY(slab_guard_violations(s, &bad));
    if (!rc && bad) {
      rc = set_err(ODE_B200_E_INVALID, "guard check: %lld doubles behind the slab's ghost cells were written", bad);
      for (int i = 0; i < 2; i++) {  // (leave the cached session clean)
        cudaMemsetAsync(s->Y[i] + s->n + 2 * s->mi.H, 0, (8 + LARGE_THREADS * LARGE_P) * sizeof(double), st);
        cudaMemsetAsync(s->K[i] + s->n + 2 * s->mi.H, 0, (8 + LARGE_THREADS * LARGE_P) * sizeof(double), st);
      }
      cudaStreamSynchronize(st);
    }
  }
  if (!rc && stats) {
    const float ms = run_ms;
    stats->n_accept = (int64_t)poll[2];
    stats->n_reject = (int64_t)poll[3];
    stats->n_rhs = (int64_t)poll[8];
    stats->status = (int32_t)poll[1];
    stats->t_final = poll[7];
    stats->kernel_ms = ms;
  }
  if (s) {  // per-call buffers go, the session stays cached
    s->trace = nullptr;
    s->trace_cap = 0;
    ode_b200_slab_set_output(s, nullptr, 0, nullptr);
  }
  if (dtrace) cudaFree(dtrace);
  if (dpts) cudaFree(dpts);
  if (dtimes) cudaFree(dtimes);
  if (!rc && stats && all_t && poll[12] > (double)all_cap) stats->status = ODE_B200_ST_OVERFLOW;
  if (rc) large_cache_release();
  return rc;
#undef LARGE_TRY
#undef LARGE_CUDA
}

// ---- one large system on several GPUs of this process ------------------------------------------------------------
// The reference integrates the whole state as one Vector in one process (oderk_adapt, src/runge_kutta.jl:232-369).
// Here the periodic grid is cut into n_gpus contiguous slabs (the first n % n_gpus one point longer), one per device;
// per step ATTEMPT every slab runs ONE kernel whose last thread block writes the slab's record (double-double error
// partial, max, NaN flag, edge values of ytrial / k_next) into every peer's HBM over NVLink, waits for the peers'
// records and runs the step controller on the gathered data -- every device takes the same decision from the same
// bits, and the result equals the single-GPU result bit for bit.  One host thread replays one CUDA graph per device.
struct MultiCache {
  std::vector<ode_b200_slab*> s;
  std::vector<cudaStream_t> st;
  std::vector<double*> dy;
  std::vector<int> dev;
  int method = -1, rhs = -1;
  long long n = 0;
};
static thread_local MultiCache g_multi;
static void multi_release() {
  MultiCache& M = g_multi;
  for (size_t g = 0; g < M.s.size(); g++) {
    cudaSetDevice(M.dev[g]);
    cudaDeviceSynchronize();
    if (M.dy[g]) cudaFree(M.dy[g]);
    if (M.s[g]) ode_b200_slab_free(M.s[g]);
    if (M.st[g]) cudaStreamDestroy(M.st[g]);
  }
  M = MultiCache();
}

int ode_b200_solve_large_multi(const ode_b200_opts* o, int32_t n_gpus, const int32_t* devices, int64_t n,
                               const double* y0, const double* params, double t0, double tend, double* y_out,
                               ode_b200_large_stats* stats, double* trace, int64_t trace_cap, int64_t* trace_n) {
  if (!o || !y0) return set_err(ODE_B200_E_INVALID, "null argument");
  int cnt = 0;
  if (cudaGetDeviceCount(&cnt) != cudaSuccess || cnt <= 0) {
    cudaGetLastError();
    return set_err(ODE_B200_E_NO_DEVICE, "no CUDA device available; libode_b200 has no CPU fallback");
  }
  if (n_gpus < 1 || n_gpus > PEER_MAX) return set_err(ODE_B200_E_INVALID, "n_gpus must be 1..%d", PEER_MAX);
  std::vector<int> dev(n_gpus);
  for (int g = 0; g < n_gpus; g++) {
    dev[g] = devices ? devices[g] : o->device + g;
    if (dev[g] < 0 || dev[g] >= cnt) return set_err(ODE_B200_E_INVALID, "device %d out of range (0..%d)", dev[g], cnt - 1);
    for (int q = 0; q < g; q++)
      if (dev[q] == dev[g]) return set_err(ODE_B200_E_INVALID, "device %d listed twice", dev[g]);
  }
  if (n_gpus == 1) {
    ode_b200_opts o1 = *o;
    o1.device = dev[0];
    o1.n_gpus = 1;
    return ode_b200_solve_large(&o1, n, y0, params, t0, tend, y_out, stats, trace, trace_cap, trace_n);
  }
  MultiCache& M = g_multi;
  int rc = 0;
  if (!M.s.empty() && (M.dev != dev || M.method != o->method || M.rhs != o->rhs || M.n != n)) multi_release();
  const long long base = n / n_gpus, rem = n % n_gpus;
  auto lo_of = [&](int g) { return g * base + (g < rem ? g : rem); };
  if (M.s.empty()) {
    M.s.assign(n_gpus, nullptr);
    M.st.assign(n_gpus, nullptr);
    M.dy.assign(n_gpus, nullptr);
    M.dev = dev;
    for (int g = 0; g < n_gpus && !rc; g++) {  // NVLink peer access, every pair
      if (cudaSetDevice(dev[g]) != cudaSuccess) rc = set_err(ODE_B200_E_CUDA, "cudaSetDevice(%d) failed", dev[g]);
      for (int q = 0; q < n_gpus && !rc; q++) {
        if (q == g) continue;
        int can = 0;
        cudaDeviceCanAccessPeer(&can, dev[g], dev[q]);
        if (!can) rc = set_err(ODE_B200_E_UNSUPPORTED, "device %d cannot address the memory of device %d (no peer access)", dev[g], dev[q]);
        if (!rc) {
          const cudaError_t e = cudaDeviceEnablePeerAccess(dev[q], 0);
          if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) rc = set_err(ODE_B200_E_CUDA, "cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e));
          cudaGetLastError();
        }
      }
    }
    std::vector<void*> blocks(n_gpus, nullptr);
    for (int g = 0; g < n_gpus && !rc; g++) {
      cudaSetDevice(dev[g]);
      if (cudaStreamCreateWithFlags(&M.st[g], cudaStreamNonBlocking) != cudaSuccess) rc = set_err(ODE_B200_E_CUDA, "cudaStreamCreate failed");
      ode_b200_opts og = *o;
      og.device = dev[g];
      const long long nl = lo_of(g + 1) - lo_of(g);
      if (!rc) rc = ode_b200_slab_create(&og, nl, n, g, n_gpus, params, t0, tend, M.st[g], &M.s[g]);
      if (!rc && cudaMalloc(&M.dy[g], (size_t)nl * 8) != cudaSuccess) rc = set_err(ODE_B200_E_NOMEM, "cudaMalloc failed");
      if (!rc) rc = ode_b200_slab_peer_alloc(M.s[g], nullptr, &blocks[g]);
    }
    for (int g = 0; g < n_gpus && !rc; g++) rc = ode_b200_slab_peer_connect(M.s[g], nullptr, blocks.data());
    if (rc) {
      cudaGetLastError();
      multi_release();
      return rc;
    }
    M.method = o->method;
    M.rhs = o->rhs;
    M.n = n;
  } else {
    if (o->points != ODE_B200_POINTS_ALL && o->points != ODE_B200_POINTS_FINAL && o->points != ODE_B200_POINTS_SPECIFIED)
      return set_err(ODE_B200_E_POINTS, "Unrecognized option points==%d", o->points);
    for (int g = 0; g < n_gpus && !rc; g++) {
      M.s[g]->trace = nullptr;
      M.s[g]->trace_cap = 0;
      rc = slab_configure(M.s[g], o, params, t0, tend);
    }
    if (rc) return rc;
  }
  double* dtrace = nullptr;
#define MULTI_CUDA(x)                                                                                \
  do {                                                                                               \
    if (!rc) {                                                                                       \
      cudaError_t e__ = (x);                                                                         \
      if (e__ != cudaSuccess) rc = set_err(ODE_B200_E_CUDA, "%s: %s", #x, cudaGetErrorString(e__));  \
    }                                                                                                \
  } while (0)
  for (int g = 0; g < n_gpus && !rc; g++) {
    MULTI_CUDA(cudaSetDevice(dev[g]));
    if (!rc) rc = ode_b200_slab_set_output(M.s[g], nullptr, 0, nullptr);
    if (g == 0 && trace && trace_cap > 0) {  // every slab records the same rows: take rank 0's
      MULTI_CUDA(cudaMalloc(&dtrace, (size_t)trace_cap * 32));
      if (!rc) rc = ode_b200_slab_set_trace(M.s[0], dtrace, trace_cap);
    }
    const long long lo = lo_of(g), nl = lo_of(g + 1) - lo;
    MULTI_CUDA(cudaMemcpyAsync(M.dy[g], y0 + lo, (size_t)nl * 8, cudaMemcpyHostToDevice, M.st[g]));
    if (!rc) rc = ode_b200_slab_set_state(M.s[g], M.dy[g]);
  }
  for (int g = 0; g < n_gpus && !rc; g++) {  // all uploads done before any kernel may wait for a peer
    MULTI_CUDA(cudaSetDevice(dev[g]));
    MULTI_CUDA(cudaStreamSynchronize(M.st[g]));
  }
  std::vector<double> polls((size_t)n_gpus * 16, 0.0);
  float ms = 0.f;
  if (!rc) rc = slabs_run(M.s.data(), n_gpus, reinterpret_cast<double(*)[16]>(polls.data()), &ms);
  for (int g = 0; g < n_gpus && !rc; g++) {
    MULTI_CUDA(cudaSetDevice(dev[g]));
    if (!rc) rc = ode_b200_slab_get_state(M.s[g], M.dy[g]);
    const long long lo = lo_of(g), nl = lo_of(g + 1) - lo;
    if (y_out) MULTI_CUDA(cudaMemcpyAsync(y_out + lo, M.dy[g], (size_t)nl * 8, cudaMemcpyDeviceToHost, M.st[g]));  // disjoint slices
  }
  if (trace && dtrace && !rc) {
    MULTI_CUDA(cudaSetDevice(dev[0]));
    const long long tn = (long long)polls[10];
    const long long m = tn < trace_cap ? tn : trace_cap;
    MULTI_CUDA(cudaMemcpyAsync(trace, dtrace, (size_t)m * 32, cudaMemcpyDeviceToHost, M.st[0]));
    if (trace_n) *trace_n = tn;
  }
  for (int g = 0; g < n_gpus && !rc; g++) {
    MULTI_CUDA(cudaSetDevice(dev[g]));
    MULTI_CUDA(cudaStreamSynchronize(M.st[g]));
  }
  if (!rc && stats) {
    const double* p0 = polls.data();
    stats->n_accept = (int64_t)p0[2];
    stats->n_reject = (int64_t)p0[3];
    stats->n_rhs = (int64_t)p0[8];
    stats->status = (int32_t)p0[1];
    stats->t_final = p0[7];
    stats->kernel_ms = ms;
    for (int g = 1; g < n_gpus; g++)  // a slab that lost its peers stops with its own status
      if ((int32_t)polls[(size_t)g * 16 + 1] != stats->status) stats->status = ODE_B200_ST_EXCHANGE;
  }
  if (dtrace) {
    cudaSetDevice(dev[0]);
    M.s[0]->trace = nullptr;
    M.s[0]->trace_cap = 0;
    cudaFree(dtrace);
  }
  if (rc) multi_release();
  return rc;
#undef MULTI_CUDA
}

}  // extern "C"

static void large_cache_release() {
  multi_release();
  LargeCache& cc = g_large_cache;
  if (cc.device >= 0) cudaSetDevice(cc.device);
  if (cc.dy) cudaFree(cc.dy);
  if (cc.s) ode_b200_slab_free(cc.s);
  if (cc.st) cudaStreamDestroy(cc.st);
  cc = LargeCache();
}

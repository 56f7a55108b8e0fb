// capi.cu -- the C ABI of libode_b200.so (include/ode_b200.h): argument checks
// with the reference's error behaviour, device memory, launches, result handles.
//
// There is no CPU path in this file: every solve entry point needs a CUDA
// device and fails with ODE_B200_E_NO_DEVICE otherwise.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "dispatch.cuh"
#include "host_util.cuh"
#include "rtc.cuh"
#include "tableaus.cuh"

namespace odeb200 {

thread_local char g_err[512] = "";
thread_local double g_last_kernel_ms = 0.0;
thread_local long long g_launches = 0;
thread_local cudaEvent_t g_ev0 = nullptr, g_ev1 = nullptr;
thread_local int g_ev_dev = -1;
thread_local bool g_ev_pending = false;

int set_err(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
void note_launch() { g_launches++; }

// ---- per-thread device arena (grow-only) ------------------------------------
struct Arena {
  int dev = -1;
  char* base = nullptr;
  size_t cap = 0, off = 0;
  cudaStream_t stream = nullptr, stream2 = nullptr;
  cudaEvent_t ev_in = nullptr, ev_done2 = nullptr;
  int ensure(int device, size_t bytes) {
    if (dev != device || cap < bytes) {
      if (base) {
        cudaSetDevice(dev);
        cudaFree(base);
        base = nullptr;
        cap = 0;
      }
      if (dev != device && stream) {  // streams and events belong to the device they were created on
        if (dev >= 0) cudaSetDevice(dev);
        cudaStreamDestroy(stream);
        cudaStreamDestroy(stream2);
        cudaEventDestroy(ev_in);
        cudaEventDestroy(ev_done2);
        stream = stream2 = nullptr;
        ev_in = ev_done2 = nullptr;
      }
      ODEB200_CUDA(cudaSetDevice(device));
      size_t want = bytes + (bytes >> 3) + (1 << 20);
      ODEB200_CUDA(cudaMalloc(&base, want));
      cap = want;
      dev = device;
    }
    if (!stream) {
      ODEB200_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
      ODEB200_CUDA(cudaStreamCreateWithFlags(&stream2, cudaStreamNonBlocking));
      ODEB200_CUDA(cudaEventCreateWithFlags(&ev_in, cudaEventDisableTiming));
      ODEB200_CUDA(cudaEventCreateWithFlags(&ev_done2, cudaEventDisableTiming));
    }
    off = 0;
    return 0;
  }
  template <class T>
  T* take(size_t n) {
    size_t bytes = (n * sizeof(T) + 255) & ~(size_t)255;
    T* p = reinterpret_cast<T*>(base + off);
    off += bytes;
    return p;
  }
};
thread_local Arena g_arena;
// Scratch of the device-pointer entry point (ode_b200_solve_ensemble_dev) -- work counters, the status buffer the two
// passes talk through, the hinit pre-pass output -- is allocated PER CALL from the device's stream-ordered memory
// pool (cudaMallocAsync on the caller's stream, freed in stream order after the last kernel), so any number of
// calls may be in flight on different streams.  The pool keeps freed blocks (release threshold = max), so a
// repeated call costs no driver allocation.
static int pool_keep(int device) {
  static bool done[64] = {};
  if (device < 0 || device >= 64 || done[device]) return 0;
  cudaMemPool_t pool = nullptr;
  ODEB200_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
  unsigned long long keep = ~0ULL;
  ODEB200_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
  done[device] = true;
  return 0;
}
// page-locked staging block of the small-call path (solve_ensemble_small)
struct SmallStage {
  char* host = nullptr;  // cudaHostAlloc, portable + mapped
  char* dev = nullptr;   // the same block as the device sees it
  size_t cap = 0;
  int ensure(size_t bytes) {
    if (cap >= bytes) return 0;
    if (host) cudaFreeHost(host);
    host = dev = nullptr;
    cap = 0;
    size_t want = bytes < (256u << 10) ? (256u << 10) : bytes;
    ODEB200_CUDA(cudaHostAlloc((void**)&host, want, cudaHostAllocPortable | cudaHostAllocMapped));
    ODEB200_CUDA(cudaHostGetDevicePointer((void**)&dev, host, 0));
    cap = want;
    return 0;
  }
  void release() {
    if (host) cudaFreeHost(host);
    host = dev = nullptr;
    cap = 0;
  }
};
thread_local SmallStage g_small;
static const size_t kSmallCallBytes = 192u << 10;
void arena_release() {  // ode_b200_trim
  if (g_arena.dev >= 0) {
    cudaSetDevice(g_arena.dev);
    if (g_arena.base) cudaFree(g_arena.base);
    if (g_arena.stream) {
      cudaStreamDestroy(g_arena.stream);
      cudaStreamDestroy(g_arena.stream2);
      cudaEventDestroy(g_arena.ev_in);
      cudaEventDestroy(g_arena.ev_done2);
    }
    g_arena = Arena();
  }
  g_small.release();
}
static size_t pad256(size_t b) { return (b + 255) & ~(size_t)255; }

// ---- names -------------------------------------------------------------------
struct NameId {
  const char* name;
  int id;
};
static const NameId kMethods[] = {
    {"ode1", ODE_B200_M_FEULER},       {"feuler", ODE_B200_M_FEULER},     {"ode2_midpoint", ODE_B200_M_MIDPOINT},
    {"midpoint", ODE_B200_M_MIDPOINT}, {"ode2_heun", ODE_B200_M_HEUN},    {"heun", ODE_B200_M_HEUN},
    {"ode4", ODE_B200_M_RK4},          {"rk4", ODE_B200_M_RK4},           {"ode21", ODE_B200_M_RK21},
    {"rk21", ODE_B200_M_RK21},         {"ode23", ODE_B200_M_RK23},        {"rk23", ODE_B200_M_RK23},
    {"ode45_fe", ODE_B200_M_RK45_FE},  {"rk45", ODE_B200_M_RK45_FE},      {"ode45", ODE_B200_M_DOPRI5},
    {"ode45_dp", ODE_B200_M_DOPRI5},   {"dopri5", ODE_B200_M_DOPRI5},     {"ode78", ODE_B200_M_FEH78},
    {"feh78", ODE_B200_M_FEH78},       {"ode23s", ODE_B200_M_ODE23S},     {"ode4s", ODE_B200_M_ODE4S_S},
    {"ode4s_s", ODE_B200_M_ODE4S_S},   {"ode4s_kr", ODE_B200_M_ODE4S_KR}, {"ode4ms", ODE_B200_M_ODE4MS},
    {"ode5ms", ODE_B200_M_ODE5MS},     {"ode_ms1", ODE_B200_M_ODE_MS1},   {"ode_ms2", ODE_B200_M_ODE_MS2},
    {"ode_ms3", ODE_B200_M_ODE_MS3},   {"ode_ms4", ODE_B200_M_ODE4MS},    {"ode_ms5", ODE_B200_M_ODE5MS},
    {"ode45_ck", ODE_B200_M_CK45},     {"ck45", ODE_B200_M_CK45}};
static const NameId kRhs[] = {{"linear", ODE_B200_RHS_LINEAR},
                              {"const", ODE_B200_RHS_CONST},
                              {"timelin", ODE_B200_RHS_TIMELIN},
                              {"rotation", ODE_B200_RHS_ROTATION},
                              {"lorenz", ODE_B200_RHS_LORENZ},
                              {"vdp", ODE_B200_RHS_VDP},
                              {"robertson", ODE_B200_RHS_ROBERTSON},
                              {"kepler", ODE_B200_RHS_KEPLER},
                              {"reaction_diffusion", ODE_B200_RHS_REACTION_DIFFUSION},
                              {"reaction_diffusion_field", ODE_B200_RHS_RD_FIELD}};
static const int kRhsDof[ODE_B200_RHS_COUNT] = {1, 1, 1, 2, 3, 2, 3, 4, 0, 0};
static const int kRhsNp[ODE_B200_RHS_COUNT] = {1, 1, 1, 0, 3, 1, 3, 0, 2, 2};

template <class T>
static double tab_get(int which, int i, int j) {
  if (i < 0 || j < 0) return NAN;
  // Rational -> Float64 exactly as the reference converts (src/ODE.jl:163)
  if (which == 0) return (i < T::S && j < T::S) ? (double)T::A[i][j].n / (double)T::A[i][j].d : NAN;
  if (which == 1) return (i < T::NB && j < T::S) ? (double)T::B[i][j].n / (double)T::B[i][j].d : NAN;
  return i < T::S ? (double)T::C[i].n / (double)T::C[i].d : NAN;
}
#define ODEB200_TAB_SWITCH(method, X)             \
  switch (method) {                               \
    case ODE_B200_M_FEULER: X(TabFEuler);         \
    case ODE_B200_M_MIDPOINT: X(TabMidpoint);     \
    case ODE_B200_M_HEUN: X(TabHeun);             \
    case ODE_B200_M_RK4: X(TabRK4);               \
    case ODE_B200_M_RK21: X(TabRK21);             \
    case ODE_B200_M_RK23: X(TabRK23);             \
    case ODE_B200_M_RK45_FE: X(TabRK45);          \
    case ODE_B200_M_DOPRI5: X(TabDopri5);         \
    case ODE_B200_M_FEH78: X(TabFeh78);           \
    case ODE_B200_M_CK45: X(TabCK45);             \
    default: break;                               \
  }

// dof / parameter count of a built-in or run-time registered rhs
static bool rhs_shape(int rhs, int* dof, int* np, bool* has_jac) {
  if (rhs >= 0 && rhs < ODE_B200_RHS_COUNT) {
    *dof = kRhsDof[rhs];
    *np = kRhsNp[rhs];
    if (has_jac) *has_jac = true;
    return true;
  }
  return user_rhs_info(rhs, dof, np, has_jac);
}

// order of ode_ms for a method id (0: not a multistep method)
int ms_order(int method) {
  switch (method) {
    case ODE_B200_M_ODE_MS1: return 1;
    case ODE_B200_M_ODE_MS2: return 2;
    case ODE_B200_M_ODE_MS3: return 3;
    case ODE_B200_M_ODE4MS: return 4;
    case ODE_B200_M_ODE5MS: return 5;
    default: return 0;
  }
}
// b[order][j] of ode_ms (0-based here).  Rows 1-4: ms_coefficients4 (src/ODE.jl:440-443) -- Julia evaluates the
// literals -1/2, 5/12, ... as Float64 divisions, so do we.  Row 5 (src/ODE.jl:183-194): for j = 0..4
//   p     = poly(roots -[0:j-1; j+1:4])          monic, from its roots by successive multiplication (exact: small integers)
//   p_int = polyint(p)                           coefficient k -> a_k/(k+1)
//   b[5, 5-j] = (-1)^j / factorial(j) / factorial(4-j) * polyval(p_int, 1)     left to right; polyval = Horner
// The exact values are [251, -1274, 2616, -2774, 1901]/720; the bits below are the ones this evaluation order gives
// (Polynomials.jl's own order cannot be checked here: no Julia -- parity of this row is unpinned).
double ms_coefficient(int order, int j) {
  static const double B4[4][4] = {{1, 0, 0, 0},
                                  {-1.0 / 2, 3.0 / 2, 0, 0},
                                  {5.0 / 12, -4.0 / 3, 23.0 / 12, 0},
                                  {-9.0 / 24, 37.0 / 24, -59.0 / 24, 55.0 / 24}};
  if (order < 1 || order > ODE_B200_MS_MAX_ORDER || j < 0 || j >= order) return NAN;
  if (order <= 4) return B4[order - 1][j];
  const int s = 5, jj = s - 1 - j;  // b[s, s - jj] (1-based) is column j (0-based)
  double p[8] = {1.0};              // ascending coefficients of the monic polynomial
  int deg = 0;
  for (int r = 0; r <= s - 1; r++) {
    if (r == jj) continue;
    const double root = -(double)r;  // multiply by (x - root)
    for (int k = deg + 1; k >= 1; k--) p[k] = p[k - 1] - root * p[k];
    p[0] = -root * p[0];
    deg++;
  }
  double pint[9] = {0.0};
  for (int k = 0; k <= deg; k++) pint[k + 1] = p[k] / (double)(k + 1);
  double v = pint[deg + 1];
  for (int k = deg; k >= 0; k--) v = pint[k] + 1.0 * v;  // polyval(p_int, 1), Horner
  double fj = 1.0, fs = 1.0;
  for (int k = 2; k <= jj; k++) fj *= k;
  for (int k = 2; k <= s - 1 - jj; k++) fs *= k;
  const double sign = (jj % 2) ? -1.0 : 1.0;
  return sign / fj / fs * v;
}

static bool is_fixed_step(int method) {
  return method <= ODE_B200_M_RK4 || method == ODE_B200_M_ODE4S_S || method == ODE_B200_M_ODE4S_KR ||
         ms_order(method) > 0;
}

static int device_ok(int dev) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0) {
    cudaGetLastError();
    return set_err(ODE_B200_E_NO_DEVICE, "no CUDA device available (%s); libode_b200 has no CPU fallback",
                   e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
  }
  if (dev < 0 || dev >= n) return set_err(ODE_B200_E_INVALID, "device %d out of range (0..%d)", dev, n - 1);
  return 0;
}

static bool fma_kernel_available(const ode_b200_opts* o);
// Checks shared by every ensemble entry point; mirrors the reference's errors.
static int validate(const ode_b200_opts* o, const ode_b200_ensemble_io* io, double t0, double tend) {
  if (!o || !io) return set_err(ODE_B200_E_INVALID, "null opts/io");
  if (o->method < 0 || o->method >= ODE_B200_M_COUNT) return set_err(ODE_B200_E_INVALID, "unknown method %d", o->method);
  int rdof = 0, rnp = 0;
  bool has_jac = true;
  if (!rhs_shape(o->rhs, &rdof, &rnp, &has_jac)) return set_err(ODE_B200_E_INVALID, "unknown rhs %d", o->rhs);
  if (o->rhs == ODE_B200_RHS_REACTION_DIFFUSION || o->rhs == ODE_B200_RHS_RD_FIELD)
    return set_err(ODE_B200_E_UNSUPPORTED, "reaction_diffusion is a large-system RHS: use ode_b200_solve_large");
  if (o->dof != rdof) return set_err(ODE_B200_E_INVALID, "rhs %d has dof %d, got %d", o->rhs, rdof, o->dof);
  if (o->n_params < rnp) return set_err(ODE_B200_E_INVALID, "rhs %d needs %d params, got %d", o->rhs, rnp, o->n_params);
  if (rnp > 0 && !io->params) return set_err(ODE_B200_E_INVALID, "params is NULL");
  if (o->params_stride != 0 && o->params_stride < rnp) return set_err(ODE_B200_E_INVALID, "params_stride < n_params");
  if (o->jacobian == 1 && !has_jac) return set_err(ODE_B200_E_INVALID, "rhs %d was registered without a Jacobian body", o->rhs);
  if (io->n_traj < 0 || (io->n_traj > 0 && !io->y0)) return set_err(ODE_B200_E_INVALID, "bad n_traj / y0");
  if (io->n_tspan < 2 || !io->tspan) return set_err(ODE_B200_E_INVALID, "tspan needs at least 2 entries");
  if (o->points != ODE_B200_POINTS_ALL && o->points != ODE_B200_POINTS_SPECIFIED && o->points != ODE_B200_POINTS_FINAL)
    return set_err(ODE_B200_E_POINTS, "Unrecognized option points==%d", o->points);  // src/runge_kutta.jl:289
  if (o->jacobian != 0 && o->jacobian != 1) return set_err(ODE_B200_E_INVALID, "jacobian must be 0 (fdjacobian) or 1 (analytic)");
  if (o->fp_mode != ODE_B200_FP_STRICT && o->fp_mode != ODE_B200_FP_FMA) return set_err(ODE_B200_E_INVALID, "fp_mode must be 0 (strict) or 1 (FMA)");
  if (o->fp_mode == ODE_B200_FP_FMA && !fma_kernel_available(o))
    return set_err(ODE_B200_E_UNSUPPORTED, "fp_mode = FMA is built for the ensemble ode45 (Dormand-Prince) kernels with a built-in "
                                           "right-hand side; everything else runs in strict mode only");
  if (!(o->reltol == o->reltol) || !(o->abstol == o->abstol) || !(o->minstep == o->minstep) ||
      !(o->maxstep == o->maxstep) || !(o->initstep == o->initstep))
    return set_err(ODE_B200_E_INVALID, "NaN in reltol/abstol/minstep/maxstep/initstep");
  const bool fixed = is_fixed_step(o->method);
  if (!fixed) {
    double tdir = tend > t0 ? 1.0 : (tend < t0 ? -1.0 : 0.0);
    if (!(tend == tend) || !(t0 == t0)) return set_err(ODE_B200_E_INVALID, "NaN in tspan");
    if (tdir == 0.0) return set_err(ODE_B200_E_ZERO_SPAN, "Zero time span");  // src/ODE.jl:88
    if (o->method != ODE_B200_M_ODE23S && o->initstep != 0.0) {
      double s = o->initstep > 0 ? 1.0 : (o->initstep < 0 ? -1.0 : 0.0);
      if (s != tdir) return set_err(ODE_B200_E_INITSTEP, "initstep has wrong sign.");  // src/runge_kutta.jl:294
    }
  }
  if ((io->pts_y || io->pts_t) && (io->pts_cap <= 0 || !io->pts_y || !io->pts_t))
    return set_err(ODE_B200_E_INVALID, "pts_t, pts_y and pts_cap must be given together");
  return 0;
}

static bool is_adaptive_rk(int method) { return method_is_adaptive_rk(method); }
// the opt-in FMA build (inst_dopri5_fma.cu): Dormand-Prince, built-in right-hand sides
extern "C" const void* odeb200_fma_dopri5_kernel(int rhs, int mode, int literal);
extern "C" const void* odeb200_fma_dopri5_init_kernel(int rhs);
static bool fma_kernel_available(const ode_b200_opts* o) {
  return o->method == ODE_B200_M_DOPRI5 && o->rhs >= 0 && o->rhs < ODE_B200_RHS_COUNT;
}
static EnsembleKernel pick_kernel_fn(const ode_b200_opts* o, int mode, int literal);

static const void* pick_kernel_builtin(const ode_b200_opts* o, int mode, int literal);

// kernel handle: a __global__ function pointer for the built-in functors, a cudaKernel_t for run-time ones
static const void* pick_kernel(const ode_b200_opts* o, int mode, int literal = 0) {
  if (o->fp_mode == ODE_B200_FP_FMA) return fma_kernel_available(o) ? odeb200_fma_dopri5_kernel(o->rhs, mode, literal) : nullptr;
  if (o->rhs >= ODE_B200_RHS_USER_BASE) return user_kernel(o->rhs, o->method, mode, literal, o->jacobian);
  return pick_kernel_builtin(o, mode, literal);
}

static const void* pick_kernel_builtin(const ode_b200_opts* o, int mode, int literal) {
  return (const void*)pick_kernel_fn(o, mode, literal);
}

static EnsembleKernel pick_kernel_fn(const ode_b200_opts* o, int mode, int literal) {
  switch (o->method) {
    case ODE_B200_M_RK21: return adapt_kernel_rk21(o->rhs, mode, literal);
    case ODE_B200_M_RK23: return adapt_kernel_rk23(o->rhs, mode, literal);
    case ODE_B200_M_RK45_FE: return adapt_kernel_rk45(o->rhs, mode, literal);
    case ODE_B200_M_DOPRI5: return adapt_kernel_dopri5(o->rhs, mode, literal);
    case ODE_B200_M_FEH78: return adapt_kernel_feh78(o->rhs, mode, literal);
    case ODE_B200_M_CK45: return adapt_kernel_ck45(o->rhs, mode, literal);
    case ODE_B200_M_ODE23S: return ode23s_kernel(o->rhs, mode, o->jacobian);
    case ODE_B200_M_ODE4S_S:
    case ODE_B200_M_ODE4S_KR: return rosenbrock_kernel(o->method, o->rhs, o->jacobian);
    default: return ms_order(o->method) > 0 ? ms_kernel(o->rhs) : fixed_kernel(o->method, o->rhs);
  }
}

// Launch on `stream` with device pointers.  `counter` points at 32 bytes of device scratch: the work
// counter of the fast pass, the redo count, and the work counter of the literal pass.  io->status must
// be non-null for the adaptive RK methods (the two passes communicate through it).
static const void* pick_init_kernel(const ode_b200_opts* o) {
  if (o->fp_mode == ODE_B200_FP_FMA) return fma_kernel_available(o) ? odeb200_fma_dopri5_init_kernel(o->rhs) : nullptr;
  if (o->rhs >= ODE_B200_RHS_USER_BASE) return user_init_kernel(o->rhs, o->method);
  switch (o->method) {
    case ODE_B200_M_RK21: return (const void*)adapt_kernel_rk21_init(o->rhs);
    case ODE_B200_M_RK23: return (const void*)adapt_kernel_rk23_init(o->rhs);
    case ODE_B200_M_RK45_FE: return (const void*)adapt_kernel_rk45_init(o->rhs);
    case ODE_B200_M_DOPRI5: return (const void*)adapt_kernel_dopri5_init(o->rhs);
    case ODE_B200_M_FEH78: return (const void*)adapt_kernel_feh78_init(o->rhs);
    case ODE_B200_M_CK45: return (const void*)adapt_kernel_ck45_init(o->rhs);
    case ODE_B200_M_ODE23S: return (const void*)ode23s_init_kernel(o->rhs, o->jacobian);
    default: return nullptr;
  }
}
// streaming pre-pass (ensemble_adapt.cuh): reads y0 from page-locked host memory while the persistent kernel runs
static const void* pick_stream_init_kernel(const ode_b200_opts* o) {
  if (o->fp_mode == ODE_B200_FP_FMA || o->rhs >= ODE_B200_RHS_USER_BASE) return nullptr;
  switch (o->method) {
    case ODE_B200_M_RK21: return adapt_kernel_rk21_stream_init(o->rhs);
    case ODE_B200_M_RK23: return adapt_kernel_rk23_stream_init(o->rhs);
    case ODE_B200_M_RK45_FE: return adapt_kernel_rk45_stream_init(o->rhs);
    case ODE_B200_M_DOPRI5: return adapt_kernel_dopri5_stream_init(o->rhs);
    case ODE_B200_M_FEH78: return adapt_kernel_feh78_stream_init(o->rhs);
    case ODE_B200_M_CK45: return adapt_kernel_ck45_stream_init(o->rhs);
    default: return nullptr;
  }
}
struct ChunkBack {         // progressive copy-back (common.cuh, EnsembleArgs::chunk_done)
  int* done = nullptr;     // device counters, one per chunk
  int* flag_dev = nullptr; // the host flags as the device sees them
  int log2 = 14;           // trajectories per chunk = 2^log2
};
struct StreamInit {        // what launch_ensemble needs for the streaming pre-pass
  int* ready = nullptr;    // one flag per STREAM_CHUNK trajectories (device)
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;

};
// doubles per trajectory the pre-pass writes: {dt0, f0[dof]} for adaptive RK, {h, F0[dof], J[dof][dof], rhs count} for ode23s
static size_t init_width(const ode_b200_opts* o) {
  return o->method == ODE_B200_M_ODE23S ? (size_t)(2 + o->dof + o->dof * o->dof) : (size_t)(o->dof + 1);
}

// ---- dispatch order: counting sort of the trajectories by the first step size (common.cuh, sched_key) -----------
// scratch words: [0] min key, [1] max key (finite keys), [2 .. 2+BINS) histogram / running offsets
// (n_dev, when given, is the element count as a device-side word: the tail hand-off's record counter)
__global__ void __launch_bounds__(256) sched_range_kernel(const double* init, int width, long long n, unsigned int* w,
                                                          const unsigned int* n_dev = nullptr) {
  __shared__ unsigned int smin[8], smax[8];
  if (n_dev) n = (long long)*n_dev;
  unsigned int lo = 0xffffffffu, hi = 0u;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    const unsigned int k = sched_key(init[i * width]);
    if (k < 0x7ff00000u) {
      lo = k < lo ? k : lo;
      hi = k > hi ? k : hi;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned int a = __shfl_down_sync(0xffffffffu, lo, o), b = __shfl_down_sync(0xffffffffu, hi, o);
    lo = a < lo ? a : lo;
    hi = b > hi ? b : hi;
  }
  if ((threadIdx.x & 31) == 0) {
    smin[threadIdx.x >> 5] = lo;
    smax[threadIdx.x >> 5] = hi;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int q = 1; q < 8; q++) {
      lo = smin[q] < lo ? smin[q] : lo;
      hi = smax[q] > hi ? smax[q] : hi;
    }
    atomicMin(&w[0], lo);
    atomicMax(&w[1], hi);
  }
}
// pass = 0: histogram; pass = 1: scatter.  Each block works on a contiguous chunk through a shared-memory histogram.
__global__ void __launch_bounds__(256) sched_bin_kernel(const double* init, int width, long long n, unsigned int* w, int* order,
                                                        int pass, const unsigned int* n_dev = nullptr) {
  __shared__ unsigned int cnt[SCHED_BINS], base[SCHED_BINS];
  if (n_dev) n = (long long)*n_dev;
  const unsigned int lo = w[0], hi = w[1];
  unsigned int* hist = w + 2;
  const long long per = (n + gridDim.x - 1) / gridDim.x, i0 = blockIdx.x * per, i1 = (i0 + per < n) ? i0 + per : n;
  for (int q = threadIdx.x; q < SCHED_BINS; q += 256) cnt[q] = 0u;
  __syncthreads();
  for (long long i = i0 + threadIdx.x; i < i1; i += 256) atomicAdd(&cnt[sched_bin(sched_key(init[i * width]), lo, hi)], 1u);
  __syncthreads();
  for (int q = threadIdx.x; q < SCHED_BINS; q += 256) {
    const unsigned int c = cnt[q];
    if (c) base[q] = atomicAdd(&hist[q], c);  // pass 0: counts; pass 1: hist holds the exclusive scan, this reserves a range
    cnt[q] = 0u;
  }
  if (pass == 0) return;
  __syncthreads();
  for (long long i = i0 + threadIdx.x; i < i1; i += 256) {
    const int bn = sched_bin(sched_key(init[i * width]), lo, hi);
    order[base[bn] + atomicAdd(&cnt[bn], 1u)] = (int)i;
  }
}
__global__ void __launch_bounds__(SCHED_BINS) sched_scan_kernel(unsigned int* w) {  // exclusive scan of the histogram, in place
  __shared__ unsigned int part[SCHED_BINS];
  unsigned int* hist = w + 2;
  const unsigned int mine = hist[threadIdx.x];
  part[threadIdx.x] = mine;
  __syncthreads();
  for (int off = 1; off < SCHED_BINS; off <<= 1) {
    const unsigned int v = threadIdx.x >= off ? part[threadIdx.x - off] : 0u;
    __syncthreads();
    part[threadIdx.x] += v;
    __syncthreads();
  }
  hist[threadIdx.x] = part[threadIdx.x] - mine;
}
// bytes of scratch for the order (index array + histogram), after the pre-pass output; 0 = natural order
static size_t order_scratch_bytes(const ode_b200_opts* o, const ode_b200_ensemble_io* io) {
  if (io->n_traj >= (1LL << 31)) return 0;
  // The tail the order removes is about half a trajectory per lane: it matters with a handful of trajectories per
  // lane and is noise with many (2^20 Lorenz trajectories = 11 per lane: the sort costs more than it saves).
  // ODE_B200_ORDER=0/1 overrides, for A/B measurements.
  if (const char* e = getenv("ODE_B200_ORDER")) {
    if (atoi(e) == 0) return 0;
  } else {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (io->n_traj > 6LL * sms * 640) return 0;  // (~640 resident lanes per SM)
  }
  return pad256((size_t)io->n_traj * sizeof(int)) + pad256((size_t)(SCHED_BINS + 2) * sizeof(unsigned int));
}
// Tail hand-off of the adaptive RK fast kernels (ensemble_adapt.cuh): scratch for one record per resident lane, their
// order and a histogram; 0 = not used for this call.  Worth it only when the lanes are kept busy by dynamic fetch
// first (several trajectories per lane); ODE_B200_DRAIN=0/1 overrides, for A/B measurements.
static size_t drain_lanes(const ode_b200_opts* o) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return (size_t)sms * 2048;  // upper bound of resident threads
}
static size_t drain_scratch_bytes(const ode_b200_opts* o, const ode_b200_ensemble_io* io) {
  if (!ODEB200_DRAIN || !is_adaptive_rk(o->method) || o->fp_mode == ODE_B200_FP_FMA || o->rhs >= ODE_B200_RHS_USER_BASE) return 0;
  const bool want_points = io->pts_y != nullptr || io->trace != nullptr;
  if ((want_points && o->points != ODE_B200_POINTS_FINAL) || io->trace) return 0;
  const size_t lanes = drain_lanes(o);
  if (const char* e = getenv("ODE_B200_DRAIN")) {
    if (atoi(e) == 0) return 0;
  } else if ((size_t)io->n_traj < 2 * lanes / 3) {  // (~640 of the 2048 lanes per SM are resident: fewer than 2 trajectories per lane)
    return 0;
  }
  return 256 + pad256(lanes * dump_width(o->dof) * sizeof(double)) + pad256(lanes * sizeof(int)) +
         pad256((size_t)(SCHED_BINS + 2) * sizeof(unsigned int));
}
// doubles of pre-pass scratch launch_ensemble can use for this call (0: hinit stays inline)
static size_t init_scratch_doubles(const ode_b200_opts* o, const ode_b200_ensemble_io* io) {
  const bool want_points = io->pts_y != nullptr || io->trace != nullptr;
  const bool mode1 = (want_points && o->points != ODE_B200_POINTS_FINAL) || io->trace;
  if (mode1 || io->n_traj < 4096 || !pick_init_kernel(o)) return 0;
  return (size_t)io->n_traj * init_width(o);
}

static int launch_ensemble(const ode_b200_opts* o, const ode_b200_ensemble_io* io, unsigned long long* counter,
                           cudaStream_t stream, bool time_it, double* init_scratch = nullptr,
                           const double* uparams_host = nullptr, void* order_scratch = nullptr, void* drain_scratch = nullptr,
                           const StreamInit* si = nullptr, int* status_out = nullptr, const ChunkBack* cb = nullptr) {
  const bool want_points = io->pts_y != nullptr || io->trace != nullptr;
  const int mode = (want_points && o->points != ODE_B200_POINTS_FINAL) || io->trace ? 1 : 0;
  // one parameter vector for the whole ensemble: the mode-0 fast kernel reads it from the constant bank
  int np_u = 0, dof_u = 0;
  rhs_shape(o->rhs, &dof_u, &np_u, nullptr);
  const bool uniform = mode == 0 && is_adaptive_rk(o->method) && o->params_stride == 0 && np_u <= 16 && (np_u == 0 || uparams_host);
  const void* k = pick_kernel(o, mode, uniform ? 2 : 0);
  if (!k) {
    if (o->rhs >= ODE_B200_RHS_USER_BASE) return ODE_B200_E_INVALID;  // last_error set by the run-time compiler
    return set_err(ODE_B200_E_UNSUPPORTED, "no kernel for method %d / rhs %d", o->method, o->rhs);
  }
  EnsembleArgs A;
  memset(&A, 0, sizeof(A));
  A.n_traj = io->n_traj;
  A.y0 = io->y0;
  A.params = io->params;
  A.tspan = io->tspan;
  A.n_tspan = io->n_tspan;
  A.params_stride = o->params_stride;
  A.points = o->points;
  A.reltol = o->reltol;
  A.abstol = o->abstol;
  A.minstep = o->minstep;
  A.maxstep = o->maxstep;
  A.initstep = o->initstep;
  A.max_steps = o->max_steps > 0 ? o->max_steps : 100000000LL;
  A.t_final = io->t_final;
  A.y_final = io->y_final;
  A.n_accept = io->n_accept;
  A.n_reject = io->n_reject;
  A.n_rhs = io->n_rhs;
  A.status = io->status;
  A.pts_t = io->pts_t;
  A.pts_y = io->pts_y;
  A.pts_n = io->pts_n;
  A.pts_cap = io->pts_y ? io->pts_cap : 0;
  A.trace_traj = io->trace ? io->trace_traj : -1;
  A.trace = io->trace;
  A.trace_cap = io->trace ? io->trace_cap : 0;
  A.trace_n = (long long*)io->trace_n;
  A.work_counter = counter;
  A.redo_count = reinterpret_cast<int*>(counter + 1);
  A.status_out = status_out;
  if (cb) {
    A.chunk_done = cb->done;
    A.chunk_flag = cb->flag_dev;
    A.chunk_log2 = cb->log2;
  }
  if (uniform)
    for (int q = 0; q < np_u; q++) A.uparams[q] = uparams_host[q];
  if (ms_order(o->method) > 0) {  // ode_ms: order and b[so-1][j] for the kernel
    A.ms_order = ms_order(o->method);
    for (int so = 1; so <= A.ms_order; so++)
      for (int j = 0; j < so; j++) A.tab[(so - 1) * ODE_B200_MS_MAX_ORDER + j] = ms_coefficient(so, j);
  }
  if (is_adaptive_rk(o->method)) {  // the tableau for the kernel-parameter constant bank (common.cuh, EnsembleArgs::tab)
    const int S = ode_b200_method_stages(o->method);
    for (int s1 = 1; s1 < S; s1++)
      for (int ss = 0; ss < s1; ss++) A.tab[tab_a_index(s1, ss)] = ode_b200_tableau(o->method, 0, s1, ss);
    for (int r = 0; r < 2; r++)
      for (int j = 0; j < S; j++) A.tab[tab_b_index(S, r, j)] = ode_b200_tableau(o->method, 1, r, j);
    for (int i = 0; i < S; i++) A.tab[tab_c_index(S, i)] = ode_b200_tableau(o->method, 2, i, 0);
  }
  const bool two_pass = is_adaptive_rk(o->method);
  if (two_pass && !A.status) return set_err(ODE_B200_E_INVALID, "internal: adaptive RK launch without a status buffer");

  int dev = 0, sms = 0, per_sm = 0;
  ODEB200_CUDA(cudaGetDevice(&dev));
  ODEB200_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int block = o->method == ODE_B200_M_ODE23S ? S23_BLOCK : ENS_BLOCK;  // threads per block of the main kernel
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, block, 0) != cudaSuccess || per_sm < 1) {
    cudaGetLastError();
    per_sm = 4;
  }
  long long need = (io->n_traj + block - 1) / block;
  long long grid = (long long)sms * per_sm;  // persistent: every lane pulls work until the counter runs out
  if (need < grid) grid = need;
  if (grid < 1) grid = 1;
  ODEB200_CUDA(cudaMemsetAsync(counter, 0, 4 * sizeof(unsigned long long), stream));
  if (time_it) {
    if (g_ev_dev != dev) {
      if (g_ev0) {
        cudaEventDestroy(g_ev0);
        cudaEventDestroy(g_ev1);
      }
      ODEB200_CUDA(cudaEventCreate(&g_ev0));
      ODEB200_CUDA(cudaEventCreate(&g_ev1));
      g_ev_dev = dev;
    }
    ODEB200_CUDA(cudaEventRecord(g_ev0, stream));
  }
  void* kargs[] = {(void*)&A};
  if (si && uniform && init_scratch) {
    // streaming pre-pass: A.y0 is the caller's page-locked buffer as the device sees it; a few blocks on the side
    // stream read it chunk by chunk while the persistent kernel (launched right after, on `stream`) already integrates
    const void* ks = pick_stream_init_kernel(o);
    if (!ks) return set_err(ODE_B200_E_UNSUPPORTED, "no streaming pre-pass for method %d / rhs %d", o->method, o->rhs);
    long long cend = (io->n_traj + STREAM_CHUNK - 1) / STREAM_CHUNK;
    ODEB200_CUDA(cudaMemsetAsync(si->ready, 0, (size_t)cend * sizeof(int), stream));
    ODEB200_CUDA(cudaEventRecord(si->ev_fork, stream));
    ODEB200_CUDA(cudaStreamWaitEvent(si->side, si->ev_fork, 0));
    A.init = init_scratch;
    int* ready = si->ready;
    // one block per chunk for the rows every resident lane starts with (all in flight at once, gone after one PCIe
    // round trip); `nlong` of them stay for the rest: every SM hosting one gives up resident CTAs of the persistent grid
    long long cfirst = (grid * block + STREAM_CHUNK - 1) / STREAM_CHUNK;
    if (getenv("ODE_B200_STREAM_ONE")) cfirst = 0;
    if (cfirst > cend) cfirst = cend;
    int nlong = 16;
    if (const char* e = getenv("ODE_B200_STREAM_GRID")) nlong = atoi(e) > 0 ? atoi(e) : nlong;
    void* sargs[] = {(void*)&A, (void*)&ready, (void*)&cfirst, (void*)&cend, (void*)&nlong};
    const long long sgrid = cfirst > nlong ? cfirst : nlong;
    ODEB200_CUDA(cudaLaunchKernel(ks, dim3((unsigned)sgrid), dim3(STREAM_CHUNK), sargs, 0, si->side));
    note_launch();
    ODEB200_CUDA(cudaEventRecord(si->ev_join, si->side));
    A.ready = si->ready;
  } else if (init_scratch && init_scratch_doubles(o, io) > 0) {  // hinit of every trajectory with full warps
    A.init = init_scratch;
    long long gi = (long long)sms * 8;
    if (need < gi) gi = need;
    ODEB200_CUDA(cudaLaunchKernel(pick_init_kernel(o), dim3((unsigned)gi), dim3(block), kargs, 0, stream));
    note_launch();
    if (order_scratch && order_scratch_bytes(o, io) > 0) {  // longest expected trajectory first
      int* order = reinterpret_cast<int*>(order_scratch);
      unsigned int* hist = reinterpret_cast<unsigned int*>((char*)order_scratch + pad256((size_t)io->n_traj * sizeof(int)));
      const unsigned int seed[2] = {0xffffffffu, 0u};
      ODEB200_CUDA(cudaMemsetAsync(hist, 0, (size_t)(SCHED_BINS + 2) * sizeof(unsigned int), stream));
      ODEB200_CUDA(cudaMemsetAsync(hist, 0xff, sizeof(unsigned int), stream));  // min key starts at the largest value
      (void)seed;
      const int w = (int)init_width(o);
      const unsigned gs = (unsigned)(io->n_traj / 2048 + 1 < (long long)sms * 4 ? io->n_traj / 2048 + 1 : (long long)sms * 4);
      sched_range_kernel<<<gs, 256, 0, stream>>>(init_scratch, w, io->n_traj, hist);
      sched_bin_kernel<<<gs, 256, 0, stream>>>(init_scratch, w, io->n_traj, hist, order, 0);
      sched_scan_kernel<<<1, SCHED_BINS, 0, stream>>>(hist);
      sched_bin_kernel<<<gs, 256, 0, stream>>>(init_scratch, w, io->n_traj, hist, order, 1);
      note_launch();
      for (int q = 0; q < 3; q++) note_launch();
      ODEB200_CUDA(cudaGetLastError());
      A.order = order;
    }
  }
  const bool drain = drain_scratch != nullptr && two_pass && mode == 0 && drain_scratch_bytes(o, io) > 0 &&
                     io->n_traj > grid * block;
  unsigned int* dcount = nullptr;
  int* dorder = nullptr;
  unsigned int* dhist = nullptr;
  if (drain) {  // phase A hands the tail over (ensemble_adapt.cuh)
    const size_t lanes = drain_lanes(o);
    char* b = reinterpret_cast<char*>(drain_scratch);
    dcount = reinterpret_cast<unsigned int*>(b);
    A.dump = reinterpret_cast<double*>(b + 256);
    A.dump_count = dcount;
    dorder = reinterpret_cast<int*>(b + 256 + pad256(lanes * dump_width(o->dof) * sizeof(double)));
    dhist = reinterpret_cast<unsigned int*>(reinterpret_cast<char*>(dorder) + pad256(lanes * sizeof(int)));
    ODEB200_CUDA(cudaMemsetAsync(dcount, 0, 256, stream));
    ODEB200_CUDA(cudaMemsetAsync(dhist, 0, (size_t)(SCHED_BINS + 2) * sizeof(unsigned int), stream));
    ODEB200_CUDA(cudaMemsetAsync(dhist, 0xff, sizeof(unsigned int), stream));
  }
  ODEB200_CUDA(cudaLaunchKernel(k, dim3((unsigned)grid), dim3(block), kargs, 0, stream));
  note_launch();
  if (drain) {  // phase B: the handed-over trajectories, least progress first, one warp = neighbours in progress
    const int w = dump_width(o->dof);
    const long long cap = grid * block;  // at most one record per lane of phase A
    const unsigned gs = (unsigned)(cap / 2048 + 1 < (long long)sms * 4 ? cap / 2048 + 1 : (long long)sms * 4);
    sched_range_kernel<<<gs, 256, 0, stream>>>(A.dump, w, cap, dhist, dcount);
    sched_bin_kernel<<<gs, 256, 0, stream>>>(A.dump, w, cap, dhist, dorder, 0, dcount);
    sched_scan_kernel<<<1, SCHED_BINS, 0, stream>>>(dhist);
    sched_bin_kernel<<<gs, 256, 0, stream>>>(A.dump, w, cap, dhist, dorder, 1, dcount);
    for (int q = 0; q < 4; q++) note_launch();
    ODEB200_CUDA(cudaGetLastError());
    A.resume = A.dump;
    A.resume_order = dorder;
    A.dump = nullptr;
    A.work_counter = counter + 3;
    ODEB200_CUDA(cudaLaunchKernel(k, dim3((unsigned)grid), dim3(block), kargs, 0, stream));
    note_launch();
    A.resume = nullptr;
    A.resume_order = nullptr;
    A.dump_count = nullptr;
    A.work_counter = counter;
  }
  if (A.ready) ODEB200_CUDA(cudaStreamWaitEvent(stream, si->ev_join, 0));  // (the pre-pass is long over)
  if (two_pass) {
    // literal pass: re-integrates the trajectories the fast pass marked ODE_B200_ST_REDO (non-finite trial
    // state or error), multiplying every tableau coefficient like the reference; returns at once if none.
    const void* kl = pick_kernel(o, mode, 1);
    if (!kl) {
      if (o->rhs >= ODE_B200_RHS_USER_BASE) return ODE_B200_E_INVALID;
      return set_err(ODE_B200_E_UNSUPPORTED, "no literal kernel for method %d / rhs %d", o->method, o->rhs);
    }
    int per_sm_l = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_l, kl, ENS_BLOCK, 0) != cudaSuccess) {
      cudaGetLastError();
      per_sm_l = 2;
    }
    long long grid_l = (long long)sms * (per_sm_l < 1 ? 1 : per_sm_l);
    if (need < grid_l) grid_l = need;
    if (grid_l < 1) grid_l = 1;
    A.work_counter = counter + 2;
    ODEB200_CUDA(cudaLaunchKernel(kl, dim3((unsigned)grid_l), dim3(ENS_BLOCK), kargs, 0, stream));
    note_launch();
  }
  if (time_it) {
    ODEB200_CUDA(cudaEventRecord(g_ev1, stream));
    g_ev_pending = true;
  }
  return 0;
}

}  // namespace odeb200

using namespace odeb200;

extern "C" {

int ode_b200_version(void) { return ODE_B200_ABI_VERSION; }
const char* ode_b200_last_error(void) { return g_err; }
int ode_b200_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}
int ode_b200_rhs_id(const char* name) {
  if (name)
    for (const NameId& e : kRhs)
      if (!strcmp(e.name, name)) return e.id;
  return ODE_B200_E_INVALID;
}
int ode_b200_method_id(const char* name) {
  if (name)
    for (const NameId& e : kMethods)
      if (!strcmp(e.name, name)) return e.id;
  return ODE_B200_E_INVALID;
}
int ode_b200_rhs_dof(int rhs) {
  int d = 0, n = 0;
  return rhs_shape(rhs, &d, &n, nullptr) ? d : ODE_B200_E_INVALID;
}
int ode_b200_rhs_nparams(int rhs) {
  int d = 0, n = 0;
  return rhs_shape(rhs, &d, &n, nullptr) ? n : ODE_B200_E_INVALID;
}
int ode_b200_method_stages(int method) {
  if (method == ODE_B200_M_ODE23S) return 3;
  if (method == ODE_B200_M_ODE4S_S || method == ODE_B200_M_ODE4S_KR) return 4;
  if (ms_order(method) > 0) return ms_order(method);
#define X(T) return T::S
  ODEB200_TAB_SWITCH(method, X)
#undef X
  return ODE_B200_E_INVALID;
}
int ode_b200_method_is_adaptive(int method) {
  if (method == ODE_B200_M_ODE23S) return 1;
  if (method == ODE_B200_M_ODE4S_S || method == ODE_B200_M_ODE4S_KR || ms_order(method) > 0) return 0;
#define X(T) return T::NB == 2
  ODEB200_TAB_SWITCH(method, X)
#undef X
  return ODE_B200_E_INVALID;
}
int ode_b200_method_is_fsal(int method) {
  if (method == ODE_B200_M_ODE4S_S || method == ODE_B200_M_ODE4S_KR || ms_order(method) > 0) return 0;
  if (method == ODE_B200_M_ODE23S) return 1;  // F0 = F2, src/ODE.jl:351
#define X(T)                       \
  {                                \
    constexpr bool f = T::fsal();  \
    return f;                      \
  }
  ODEB200_TAB_SWITCH(method, X)
#undef X
  return ODE_B200_E_INVALID;
}
int ode_b200_method_ms_order(int method) { return ms_order(method); }
double ode_b200_ms_coefficient(int order, int j) { return ms_coefficient(order, j - 1); }
double ode_b200_tableau(int method, int which, int i, int j) {
#define X(T) return tab_get<T>(which, i, j)
  ODEB200_TAB_SWITCH(method, X)
#undef X
  return NAN;
}

void* ode_b200_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable | cudaHostAllocMapped) != cudaSuccess) {
    set_err(ODE_B200_E_NOMEM, "cudaHostAlloc(%zu) failed: %s", bytes, cudaGetErrorString(cudaGetLastError()));
    return nullptr;
  }
  return p;
}
void ode_b200_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

double ode_b200_last_kernel_ms(void) {
  if (g_ev_pending) {
    float ms = 0.f;
    if (cudaEventSynchronize(g_ev1) == cudaSuccess && cudaEventElapsedTime(&ms, g_ev0, g_ev1) == cudaSuccess)
      g_last_kernel_ms = ms;
    g_ev_pending = false;
  }
  return g_last_kernel_ms;
}
int64_t ode_b200_launch_count(int reset) {
  long long v = g_launches;
  if (reset) g_launches = 0;
  return v;
}

int ode_b200_solve_ensemble_dev(const ode_b200_opts* opts, const ode_b200_ensemble_io* io, void* stream) {
  if (!opts || !io) return set_err(ODE_B200_E_INVALID, "null opts/io");
  int rc = device_ok(opts->device);
  if (rc) return rc;
  ODEB200_CUDA(cudaSetDevice(opts->device));
  // the span is needed on the host for the reference's argument errors
  double ends[2] = {0, 0};
  double upar[16];
  int np_dev = 0, dof_dev = 0;
  const bool fetch_upar = rhs_shape(opts->rhs, &dof_dev, &np_dev, nullptr) && np_dev > 0 && np_dev <= 16 &&
                          opts->params_stride == 0 && io->params != nullptr;
  if (io->n_tspan >= 2 && io->tspan) {
    if (fetch_upar)
      ODEB200_CUDA(cudaMemcpyAsync(upar, io->params, (size_t)np_dev * 8, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    ODEB200_CUDA(cudaMemcpyAsync(&ends[0], io->tspan, 8, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    ODEB200_CUDA(cudaMemcpyAsync(&ends[1], io->tspan + (io->n_tspan - 1), 8, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    ODEB200_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  }
  rc = validate(opts, io, ends[0], ends[1]);
  if (rc) return rc;
  if (io->n_traj == 0) return 0;
  rc = pool_keep(opts->device);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  ode_b200_ensemble_io d = *io;
  const bool need_status = !d.status && is_adaptive_rk(opts->method);  // the fast and literal passes talk through `status`
  const size_t status_bytes = need_status ? pad256((size_t)io->n_traj * sizeof(int)) : 0;
  const size_t want = init_scratch_doubles(opts, &d);
  const size_t order_bytes = want > 0 ? order_scratch_bytes(opts, &d) : 0;
  const size_t drain_bytes = drain_scratch_bytes(opts, &d);
  char* blk = nullptr;
  ODEB200_CUDA(cudaMallocAsync((void**)&blk, 256 + status_bytes + pad256(want * sizeof(double)) + order_bytes + drain_bytes, st));
  unsigned long long* counter = reinterpret_cast<unsigned long long*>(blk);
  if (need_status) d.status = reinterpret_cast<int*>(blk + 256);
  double* init_scratch = want > 0 ? reinterpret_cast<double*>(blk + 256 + status_bytes) : nullptr;
  void* order_scratch = order_bytes ? blk + 256 + status_bytes + pad256(want * sizeof(double)) : nullptr;
  void* drain_scratch = drain_bytes ? blk + 256 + status_bytes + pad256(want * sizeof(double)) + order_bytes : nullptr;
  rc = launch_ensemble(opts, &d, counter, st, true, init_scratch, fetch_upar ? upar : nullptr, order_scratch, drain_scratch);
  cudaFreeAsync(blk, st);  // stream-ordered: after the kernels above
  return rc;
}

// ---- small calls (a single system, a handful of trajectories) -----------------------------------
// A call that moves a few kilobytes is dominated by the number of driver calls, not by bytes: the general
// path below issues ~15 (three uploads, a memset, the launches, up to nine downloads).  Here the inputs are
// packed into one page-locked staging block and uploaded with a single copy, and the kernel writes every
// output straight into the device-mapped half of the same block; after the stream synchronises the results
// are copied out by the CPU.  One upload, one memset, the launches, one synchronise.

static int solve_ensemble_small(const ode_b200_opts* opts, const ode_b200_ensemble_io* io, size_t n_par) {
  const long long n = io->n_traj;
  const int D = opts->dof;
  const bool pts = io->pts_y != nullptr;
  // layout of the staging block: inputs first (uploaded), then outputs (written by the kernel)
  size_t off = 0;
  auto place = [&](size_t bytes) {
    size_t o = off;
    off += pad256(bytes);
    return o;
  };
  const size_t o_y0 = place((size_t)n * D * 8), o_par = place(n_par * 8), o_ts = place((size_t)io->n_tspan * 8);
  const size_t in_bytes = off;
  const size_t o_tf = place(n * 8), o_yf = place((size_t)n * D * 8), o_na = place(n * 4), o_nr = place(n * 4),
               o_nf = place(n * 4), o_st = place(n * 4);
  const size_t o_pt = pts ? place((size_t)n * io->pts_cap * 8) : 0, o_py = pts ? place((size_t)n * io->pts_cap * D * 8) : 0,
               o_pn = pts ? place(n * 4) : 0;
  const size_t o_tr = io->trace ? place((size_t)io->trace_cap * 32) : 0, o_tn = io->trace ? place(8) : 0;
  int rc = g_arena.ensure(opts->device, 256 + pad256(in_bytes));  // (also makes opts->device current)
  if (rc) return rc;
  rc = g_small.ensure(off);
  if (rc) return rc;
  cudaStream_t st = g_arena.stream;
  unsigned long long* counter = g_arena.take<unsigned long long>(1);
  char* din = reinterpret_cast<char*>(g_arena.take<char>(in_bytes));
  char* h = g_small.host;
  char* dv = g_small.dev;
  memcpy(h + o_y0, io->y0, (size_t)n * D * 8);
  if (n_par) memcpy(h + o_par, io->params, n_par * 8);
  memcpy(h + o_ts, io->tspan, (size_t)io->n_tspan * 8);
  if (io->trace) memset(h + o_tn, 0, 8);
  if (pts) {  // rows a truncated integration never reaches (minstep stop with points = :specified) read as zeros
    memset(h + o_pt, 0, (size_t)n * io->pts_cap * 8);
    memset(h + o_py, 0, (size_t)n * io->pts_cap * D * 8);
  }
  ODEB200_CUDA(cudaMemcpyAsync(din, h, in_bytes, cudaMemcpyHostToDevice, st));
  ode_b200_ensemble_io d = *io;
  d.y0 = reinterpret_cast<double*>(din + o_y0);
  d.params = n_par ? reinterpret_cast<double*>(din + o_par) : nullptr;
  d.tspan = reinterpret_cast<double*>(din + o_ts);
  d.t_final = io->t_final ? reinterpret_cast<double*>(dv + o_tf) : nullptr;
  d.y_final = io->y_final ? reinterpret_cast<double*>(dv + o_yf) : nullptr;
  d.n_accept = io->n_accept ? reinterpret_cast<int*>(dv + o_na) : nullptr;
  d.n_reject = io->n_reject ? reinterpret_cast<int*>(dv + o_nr) : nullptr;
  d.n_rhs = io->n_rhs ? reinterpret_cast<int*>(dv + o_nf) : nullptr;
  d.status = reinterpret_cast<int*>(dv + o_st);
  d.pts_t = pts ? reinterpret_cast<double*>(dv + o_pt) : nullptr;
  d.pts_y = pts ? reinterpret_cast<double*>(dv + o_py) : nullptr;
  d.pts_n = (pts && io->pts_n) ? reinterpret_cast<int*>(dv + o_pn) : nullptr;
  if (io->trace) {
    d.trace = reinterpret_cast<double*>(dv + o_tr);
    d.trace_n = reinterpret_cast<int64_t*>(dv + o_tn);
  }
  rc = launch_ensemble(opts, &d, counter, st, true, nullptr, opts->params_stride == 0 ? io->params : nullptr);
  if (rc) return rc;
  ODEB200_CUDA(cudaStreamSynchronize(st));
  if (io->t_final) memcpy(io->t_final, h + o_tf, n * 8);
  if (io->y_final) memcpy(io->y_final, h + o_yf, (size_t)n * D * 8);
  if (io->n_accept) memcpy(io->n_accept, h + o_na, n * 4);
  if (io->n_reject) memcpy(io->n_reject, h + o_nr, n * 4);
  if (io->n_rhs) memcpy(io->n_rhs, h + o_nf, n * 4);
  if (io->status) memcpy(io->status, h + o_st, n * 4);
  if (pts) {
    memcpy(io->pts_t, h + o_pt, (size_t)n * io->pts_cap * 8);
    memcpy(io->pts_y, h + o_py, (size_t)n * io->pts_cap * D * 8);
    if (io->pts_n) memcpy(io->pts_n, h + o_pn, n * 4);
  }
  if (io->trace) {
    memcpy(io->trace, h + o_tr, (size_t)io->trace_cap * 32);
    if (io->trace_n) memcpy(io->trace_n, h + o_tn, 8);
  }
  return 0;
}

// ---- ensembles on several GPUs of this process (opts->n_gpus > 1) ---------------------------------------------------
// Trajectories are independent (the reference has no cross-trajectory state), so the ensemble is cut into n_gpus
// contiguous blocks; one persistent host thread per device runs the single-GPU path on its block, reading its
// slice of the caller's inputs and writing its slice of the caller's outputs.  No collective, no gather step:
// when the call returns the caller's buffers hold the whole ensemble.
namespace {
struct FanWorker {
  std::thread th;
  std::mutex mu;
  std::condition_variable cv;
  std::function<void()> job;
  bool has_job = false, stop = false;
  void loop() {
    std::unique_lock<std::mutex> lk(mu);
    while (true) {
      cv.wait(lk, [&] { return has_job || stop; });
      if (stop) return;
      std::function<void()> j = std::move(job);
      lk.unlock();
      j();
      lk.lock();
      has_job = false;
      cv.notify_all();
    }
  }
  void submit(std::function<void()> j) {
    std::unique_lock<std::mutex> lk(mu);
    job = std::move(j);
    has_job = true;
    cv.notify_all();
  }
  void wait() {
    std::unique_lock<std::mutex> lk(mu);
    cv.wait(lk, [&] { return !has_job; });
  }
};
std::mutex g_fan_mu;                      // one fan-out at a time per process (the workers' arenas are per device)
std::vector<FanWorker*> g_fan_workers;    // index = device ordinal; threads live until the process ends
FanWorker* fan_worker(int device) {
  if ((int)g_fan_workers.size() <= device) g_fan_workers.resize(device + 1, nullptr);
  if (!g_fan_workers[device]) {
    FanWorker* w = new FanWorker();
    w->th = std::thread([w] { w->loop(); });
    w->th.detach();
    g_fan_workers[device] = w;
  }
  return g_fan_workers[device];
}
}  // namespace

static int solve_ensemble_fanout(const ode_b200_opts* opts, const ode_b200_ensemble_io* io) {
  const int G = opts->n_gpus;
  int cnt = 0;
  if (cudaGetDeviceCount(&cnt) != cudaSuccess || cnt <= 0) {
    cudaGetLastError();
    return set_err(ODE_B200_E_NO_DEVICE, "no CUDA device available; libode_b200 has no CPU fallback");
  }
  if (opts->device < 0 || opts->device + G > cnt)
    return set_err(ODE_B200_E_INVALID, "n_gpus %d from device %d: only %d devices visible", G, opts->device, cnt);
  if (io->n_tspan < 2 || !io->tspan) return set_err(ODE_B200_E_INVALID, "tspan needs at least 2 entries");
  int rc = validate(opts, io, io->tspan[0], io->tspan[io->n_tspan - 1]);
  if (rc) return rc;
  const long long n = io->n_traj;
  const int D = opts->dof;
  std::lock_guard<std::mutex> fan_lock(g_fan_mu);
  std::vector<int> rcs(G, 0);
  std::vector<std::string> errs(G);
  std::vector<double> kms(G, 0.0);
  std::vector<long long> launches(G, 0);
  for (int g = 0; g < G; g++) {
    const long long lo = n * g / G, hi = n * (g + 1) / G;
    ode_b200_opts og = *opts;
    og.device = opts->device + g;
    og.n_gpus = 1;
    ode_b200_ensemble_io ig = *io;
    ig.n_traj = hi - lo;
    ig.y0 = io->y0 + lo * D;
    if (io->params && opts->params_stride) ig.params = io->params + lo * opts->params_stride;
    if (io->t_final) ig.t_final = io->t_final + lo;
    if (io->y_final) ig.y_final = io->y_final + lo * D;
    if (io->n_accept) ig.n_accept = io->n_accept + lo;
    if (io->n_reject) ig.n_reject = io->n_reject + lo;
    if (io->n_rhs) ig.n_rhs = io->n_rhs + lo;
    if (io->status) ig.status = io->status + lo;
    if (io->pts_y) {
      ig.pts_t = io->pts_t + lo * io->pts_cap;
      ig.pts_y = io->pts_y + lo * io->pts_cap * D;
      if (io->pts_n) ig.pts_n = io->pts_n + lo;
    }
    if (io->trace && io->trace_traj >= lo && io->trace_traj < hi) {
      ig.trace_traj = io->trace_traj - lo;
    } else {
      ig.trace = nullptr;
      ig.trace_n = nullptr;
      ig.trace_traj = -1;
    }
    FanWorker* w = fan_worker(og.device);
    int* prc = &rcs[g];
    std::string* perr = &errs[g];
    double* pk = &kms[g];
    long long* pl = &launches[g];
    w->submit([og, ig, prc, perr, pk, pl]() {
      ode_b200_launch_count(1);
      *prc = ig.n_traj > 0 ? ode_b200_solve_ensemble(&og, &ig) : 0;
      if (*prc) *perr = ode_b200_last_error();
      *pk = ig.n_traj > 0 ? ode_b200_last_kernel_ms() : 0.0;
      *pl = ode_b200_launch_count(1);
    });
  }
  for (int g = 0; g < G; g++) fan_worker(opts->device + g)->wait();
  double km = 0.0;
  for (int g = 0; g < G; g++) {
    km = kms[g] > km ? kms[g] : km;
    g_launches += launches[g];
    if (rcs[g] && !rc) rc = set_err(rcs[g], "device %d: %s", opts->device + g, errs[g].c_str());
  }
  g_last_kernel_ms = km;  // the slowest device's kernel
  g_ev_pending = false;
  return rc;
}

int ode_b200_solve_ensemble(const ode_b200_opts* opts, const ode_b200_ensemble_io* io) {
  if (!opts || !io) return set_err(ODE_B200_E_INVALID, "null opts/io");
  if (opts->n_gpus > 1) return solve_ensemble_fanout(opts, io);
  int rc = device_ok(opts->device);
  if (rc) return rc;
  if (io->n_tspan < 2 || !io->tspan) return set_err(ODE_B200_E_INVALID, "tspan needs at least 2 entries");
  rc = validate(opts, io, io->tspan[0], io->tspan[io->n_tspan - 1]);
  if (rc) return rc;
  const long long n = io->n_traj;
  if (n == 0) return 0;
  const int D = opts->dof;
  int np = 0, dof_unused = 0;
  rhs_shape(opts->rhs, &dof_unused, &np, nullptr);
  const size_t n_par = np ? (opts->params_stride ? (size_t)n * opts->params_stride : (size_t)np) : 0;
  const bool pts = io->pts_y != nullptr;
  size_t bytes = 256 + pad256(n * D * 8) + pad256(n_par * 8) + pad256((size_t)io->n_tspan * 8) + pad256(n * 8) +
                 pad256(n * D * 8) + 4 * pad256(n * 4);
  if (pts) bytes += pad256((size_t)n * io->pts_cap * 8) + pad256((size_t)n * io->pts_cap * D * 8) + pad256(n * 4);
  if (io->trace) bytes += pad256((size_t)io->trace_cap * 32) + 256;
  if (bytes <= kSmallCallBytes && !getenv("ODE_B200_NO_SMALL_PATH")) return solve_ensemble_small(opts, io, n_par);
  size_t init_doubles = init_scratch_doubles(opts, io);  // hinit pre-pass scratch (0: inline)
  const size_t order_bytes = init_doubles ? order_scratch_bytes(opts, io) : 0;
  const size_t drain_bytes = drain_scratch_bytes(opts, io);
  // Streaming pre-pass: with y0 in page-locked (device-mapped) host memory and one parameter vector for the whole
  // ensemble, the pre-pass reads y0 over PCIe chunk by chunk while the persistent kernel already integrates, instead
  // of a host-to-device copy in front of everything (ensemble_adapt.cuh).  ODE_B200_STREAM_INIT=0/1 overrides.
  const double* y0_mapped = nullptr;
  {
    bool want = n >= (1 << 18);
    if (const char* e = getenv("ODE_B200_STREAM_INIT")) want = atoi(e) != 0;
    cudaPointerAttributes at;
    if (want && init_doubles > 0 && order_bytes == 0 && !getenv("ODE_B200_CHUNKS") && is_adaptive_rk(opts->method) &&
        opts->params_stride == 0 && np <= 16 && !pts && !io->trace && pick_stream_init_kernel(opts) != nullptr) {
      if (cudaPointerGetAttributes(&at, io->y0) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer &&
          ((size_t)at.devicePointer & 15) == 0)  // (the pre-pass reads y0 in 16-byte pieces)
        y0_mapped = (const double*)at.devicePointer;
      else
        cudaGetLastError();
    }
  }
  // Progressive copy-back: staged (not kernel-written) outputs of a large adaptive-RK ensemble go back to the host in
  // chunks while the kernel is still running; the host thread of this blocking call does the polling.  With many GPUs
  // behind one host this beats both a copy phase after the kernel and the kernel's own small writes into host memory
  // (profiles/ensemble_kernels_r2.md).  ODE_B200_PROGRESSIVE=0/1 overrides.
  ChunkBack cback;
  bool progressive = n >= (1 << 17) && opts->host_output != 1;
  if (const char* e = getenv("ODE_B200_PROGRESSIVE")) progressive = atoi(e) != 0;
  progressive = progressive && is_adaptive_rk(opts->method) && !pts && !io->trace && !getenv("ODE_B200_CHUNKS") &&
                opts->fp_mode != ODE_B200_FP_FMA && opts->rhs < ODE_B200_RHS_USER_BASE;
  const long long n_back = progressive ? ((n + (1LL << cback.log2) - 1) >> cback.log2) : 0;
  if (progressive) bytes += pad256((size_t)n_back * sizeof(int));
  size_t ready_bytes = 0;
  if (y0_mapped) {
    init_doubles = (size_t)n * (2 * D + 1);
    ready_bytes = pad256(((size_t)n / STREAM_CHUNK + 1) * sizeof(int));
  }
  bytes += pad256(init_doubles * 8) + order_bytes + 256 + drain_bytes + 256 + ready_bytes;
  rc = g_arena.ensure(opts->device, bytes);
  if (rc) return rc;
  cudaStream_t st = g_arena.stream;
  ode_b200_ensemble_io d = *io;
  unsigned long long* counter = g_arena.take<unsigned long long>(1);
  double* dy0 = g_arena.take<double>((size_t)n * D);
  double* dpar = n_par ? g_arena.take<double>(n_par) : nullptr;
  double* dts = g_arena.take<double>(io->n_tspan);
  d.y0 = dy0;
  d.params = dpar;
  d.tspan = dts;
  // opts.host_output = 1: output buffers the caller keeps in pinned (page-locked, device-mapped) host memory are
  // written by the kernel itself as each trajectory ends, so the results cross PCIe while the integration is
  // still running instead of in a D2H phase after it.  Pageable buffers are staged in HBM and copied back.
  // (`status` is always staged: the fast and literal passes talk through it.)  The environment variable
  // ODE_B200_ZEROCOPY=0/1 overrides the option, for A/B measurements.
  bool zc_on = opts->host_output == 1;
  if (const char* e = getenv("ODE_B200_ZEROCOPY")) zc_on = atoi(e) != 0;
  auto mapped = [&](const void* h) -> void* {
    if (!zc_on || !h) return nullptr;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, h) != cudaSuccess) {
      cudaGetLastError();
      return nullptr;
    }
    return at.type == cudaMemoryTypeHost ? at.devicePointer : nullptr;
  };
  double* zc_t = (double*)mapped(io->t_final);
  double* zc_y = (double*)mapped(io->y_final);
  int* zc_na = (int*)mapped(io->n_accept);
  int* zc_nr = (int*)mapped(io->n_reject);
  int* zc_nf = (int*)mapped(io->n_rhs);
  d.t_final = zc_t ? zc_t : (io->t_final ? g_arena.take<double>(n) : nullptr);
  d.y_final = zc_y ? zc_y : (io->y_final ? g_arena.take<double>((size_t)n * D) : nullptr);
  d.n_accept = zc_na ? zc_na : (io->n_accept ? g_arena.take<int>(n) : nullptr);
  d.n_reject = zc_nr ? zc_nr : (io->n_reject ? g_arena.take<int>(n) : nullptr);
  d.n_rhs = zc_nf ? zc_nf : (io->n_rhs ? g_arena.take<int>(n) : nullptr);
  // status: the fast and literal adaptive-RK passes talk through a device copy; a mapped caller buffer gets the final
  // values from the kernels directly (adaptive RK: next to the device copy; other methods write status once)
  int* zc_st = (int*)mapped(io->status);
  const bool rk_two_pass = is_adaptive_rk(opts->method);
  d.status = (zc_st && !rk_two_pass) ? zc_st : g_arena.take<int>(n);
  double* dinit = init_doubles ? g_arena.take<double>(init_doubles) : nullptr;
  char* dorder = order_bytes ? g_arena.take<char>(order_bytes) : nullptr;
  char* ddrain = drain_bytes ? g_arena.take<char>(drain_bytes) : nullptr;
  int* back_flags = nullptr;  // host side of the chunk flags
  if (progressive) {
    cback.done = g_arena.take<int>((size_t)n_back);
    rc = g_small.ensure((size_t)n_back * sizeof(int));
    if (rc) return rc;
    back_flags = reinterpret_cast<int*>(g_small.host);
    cback.flag_dev = reinterpret_cast<int*>(g_small.dev);
    memset(back_flags, 0, (size_t)n_back * sizeof(int));
    ODEB200_CUDA(cudaMemsetAsync(cback.done, 0, (size_t)n_back * sizeof(int), st));
  }
  StreamInit sinit;
  if (y0_mapped) {
    sinit.ready = reinterpret_cast<int*>(g_arena.take<char>(ready_bytes));
    sinit.side = g_arena.stream2;
    sinit.ev_fork = g_arena.ev_in;
    sinit.ev_join = g_arena.ev_done2;
    d.y0 = y0_mapped;  // nothing is uploaded: every reader takes y0 from the caller's buffer
  }
  if (pts) {
    d.pts_t = g_arena.take<double>((size_t)n * io->pts_cap);
    d.pts_y = g_arena.take<double>((size_t)n * io->pts_cap * D);
    d.pts_n = io->pts_n ? g_arena.take<int>(n) : nullptr;
    // rows a truncated integration never reaches (minstep stop with points = :specified) read as zeros
    ODEB200_CUDA(cudaMemsetAsync(d.pts_t, 0, (size_t)n * io->pts_cap * 8, st));
    ODEB200_CUDA(cudaMemsetAsync(d.pts_y, 0, (size_t)n * io->pts_cap * D * 8, st));
  } else {
    d.pts_t = d.pts_y = nullptr;
    d.pts_n = nullptr;
  }
  long long* dtrace_n = nullptr;
  if (io->trace) {
    d.trace = g_arena.take<double>((size_t)io->trace_cap * 4);
    dtrace_n = g_arena.take<long long>(1);
    d.trace_n = (int64_t*)dtrace_n;
    ODEB200_CUDA(cudaMemsetAsync(dtrace_n, 0, 8, st));
  }
  // Optional chunk pipeline (ODE_B200_CHUNKS=k): the ensemble is cut into k chunks on two streams so the
  // H2D of one chunk and the D2H of another overlap a running kernel.  Off by default: measured on B200 it
  // only pays for very large, cheap-per-trajectory ensembles (2^20 Lorenz/ode45: 12.99 ms -> 12.46 ms with
  // 16 chunks) and costs 6-200 % elsewhere, because every chunk's persistent kernel has its own
  // lane-starved tail.  Results do not depend on the chunking.
  int n_chunks = 1;
  if (const char* e = getenv("ODE_B200_CHUNKS")) {
    int v = atoi(e);
    if (v >= 1 && v <= 64 && !pts && !io->trace && n >= v) n_chunks = v;
  }
  unsigned long long* counters = counter;
  if (n_chunks > 1) counters = g_arena.take<unsigned long long>(32 * n_chunks);  // 256 B apart, 32 B used each
  if (n_par && !opts->params_stride) ODEB200_CUDA(cudaMemcpyAsync(dpar, io->params, n_par * 8, cudaMemcpyHostToDevice, st));
  ODEB200_CUDA(cudaMemcpyAsync(dts, io->tspan, (size_t)io->n_tspan * 8, cudaMemcpyHostToDevice, st));
  if (g_ev_dev != opts->device) {
    if (g_ev0) {
      cudaEventDestroy(g_ev0);
      cudaEventDestroy(g_ev1);
    }
    ODEB200_CUDA(cudaEventCreate(&g_ev0));
    ODEB200_CUDA(cudaEventCreate(&g_ev1));
    g_ev_dev = opts->device;
  }
  // ode_b200_last_kernel_ms: kernel only when there is a single chunk, else the device span of the
  // whole chunk pipeline (copies included)
  if (n_chunks > 1) ODEB200_CUDA(cudaEventRecord(g_ev0, st));
  ODEB200_CUDA(cudaEventRecord(g_arena.ev_in, st));
  ODEB200_CUDA(cudaStreamWaitEvent(g_arena.stream2, g_arena.ev_in, 0));
  for (int c = 0; c < n_chunks; c++) {
    const long long lo = n * c / n_chunks, hi = n * (c + 1) / n_chunks, m = hi - lo;
    cudaStream_t cs = (c & 1) ? g_arena.stream2 : st;
    ode_b200_ensemble_io dc = d;
    dc.n_traj = m;
    if (!y0_mapped) {
      dc.y0 = dy0 + lo * D;
      ODEB200_CUDA(cudaMemcpyAsync(dy0 + lo * D, io->y0 + lo * D, (size_t)m * D * 8, cudaMemcpyHostToDevice, cs));
    }
    if (n_par && opts->params_stride) {
      dc.params = dpar + lo * opts->params_stride;
      ODEB200_CUDA(cudaMemcpyAsync(dpar + lo * opts->params_stride, io->params + lo * opts->params_stride,
                                   (size_t)m * opts->params_stride * 8, cudaMemcpyHostToDevice, cs));
    }
    if (d.t_final) dc.t_final = d.t_final + lo;
    if (d.y_final) dc.y_final = d.y_final + lo * D;
    if (d.n_accept) dc.n_accept = d.n_accept + lo;
    if (d.n_reject) dc.n_reject = d.n_reject + lo;
    if (d.n_rhs) dc.n_rhs = d.n_rhs + lo;
    if (d.status) dc.status = d.status + lo;
    if (pts) {
      dc.pts_t = d.pts_t + lo * io->pts_cap;
      dc.pts_y = d.pts_y + lo * io->pts_cap * D;
      if (d.pts_n) dc.pts_n = d.pts_n + lo;
    }
    if (n_chunks == 1) ODEB200_CUDA(cudaEventRecord(g_ev0, cs));
    rc = launch_ensemble(opts, &dc, n_chunks > 1 ? counters + 32 * c : counter, cs, false,
                         dinit ? dinit + lo * init_width(opts) : nullptr, opts->params_stride == 0 ? io->params : nullptr,
                         n_chunks == 1 ? (void*)dorder : nullptr, n_chunks == 1 ? (void*)ddrain : nullptr,
                         y0_mapped ? &sinit : nullptr, (zc_st && rk_two_pass) ? zc_st + lo : nullptr,
                         (progressive && n_chunks == 1) ? &cback : nullptr);
    if (rc) return rc;
    if (n_chunks == 1) ODEB200_CUDA(cudaEventRecord(g_ev1, cs));
    if (progressive && n_chunks == 1) {
      // copy every chunk whose flag is up (on the second stream: the data of a flagged chunk is final), until the
      // kernels are done; then the rest
      std::vector<char> copied((size_t)n_back, 0);
      auto copy_chunk = [&](long long c) -> cudaError_t {
        const long long a = c << cback.log2, b = (a + (1LL << cback.log2) < n) ? a + (1LL << cback.log2) : n, k = b - a;
        cudaStream_t s2 = g_arena.stream2;
        cudaError_t e = cudaSuccess;
        if (io->t_final && !zc_t && e == cudaSuccess) e = cudaMemcpyAsync(io->t_final + a, dc.t_final + a, k * 8, cudaMemcpyDeviceToHost, s2);
        if (io->y_final && !zc_y && e == cudaSuccess) e = cudaMemcpyAsync(io->y_final + a * D, dc.y_final + a * D, (size_t)k * D * 8, cudaMemcpyDeviceToHost, s2);
        if (io->n_accept && !zc_na && e == cudaSuccess) e = cudaMemcpyAsync(io->n_accept + a, dc.n_accept + a, k * 4, cudaMemcpyDeviceToHost, s2);
        if (io->n_reject && !zc_nr && e == cudaSuccess) e = cudaMemcpyAsync(io->n_reject + a, dc.n_reject + a, k * 4, cudaMemcpyDeviceToHost, s2);
        if (io->n_rhs && !zc_nf && e == cudaSuccess) e = cudaMemcpyAsync(io->n_rhs + a, dc.n_rhs + a, k * 4, cudaMemcpyDeviceToHost, s2);
        if (io->status && !zc_st && e == cudaSuccess) e = cudaMemcpyAsync(io->status + a, dc.status + a, k * 4, cudaMemcpyDeviceToHost, s2);
        return e;
      };
      long long first = 0;
      bool kernels_done = false;
      while (!kernels_done) {
        const cudaError_t q = cudaEventQuery(g_ev1);
        if (q == cudaSuccess) kernels_done = true;
        else if (q != cudaErrorNotReady) ODEB200_CUDA(q);
        while (first < n_back && copied[first]) first++;
        const long long last = first + 64 < n_back ? first + 64 : n_back;  // (chunks complete roughly in order)
        for (long long c = first; c < last; c++) {
          if (!copied[c] && *(volatile int*)&back_flags[c] != 0) {
            ODEB200_CUDA(copy_chunk(c));
            copied[c] = 1;
          }
        }
      }
      ODEB200_CUDA(cudaStreamWaitEvent(g_arena.stream2, g_ev1, 0));
      for (long long c = 0; c < n_back; c++)
        if (!copied[c]) ODEB200_CUDA(copy_chunk(c));
      continue;  // (everything is on its way: skip the copy phase below)
    }
    if (io->t_final && !zc_t) ODEB200_CUDA(cudaMemcpyAsync(io->t_final + lo, dc.t_final, m * 8, cudaMemcpyDeviceToHost, cs));
    if (io->y_final && !zc_y) ODEB200_CUDA(cudaMemcpyAsync(io->y_final + lo * D, dc.y_final, (size_t)m * D * 8, cudaMemcpyDeviceToHost, cs));
    if (io->n_accept && !zc_na) ODEB200_CUDA(cudaMemcpyAsync(io->n_accept + lo, dc.n_accept, m * 4, cudaMemcpyDeviceToHost, cs));
    if (io->n_reject && !zc_nr) ODEB200_CUDA(cudaMemcpyAsync(io->n_reject + lo, dc.n_reject, m * 4, cudaMemcpyDeviceToHost, cs));
    if (io->n_rhs && !zc_nf) ODEB200_CUDA(cudaMemcpyAsync(io->n_rhs + lo, dc.n_rhs, m * 4, cudaMemcpyDeviceToHost, cs));
    if (io->status && !zc_st) ODEB200_CUDA(cudaMemcpyAsync(io->status + lo, dc.status, m * 4, cudaMemcpyDeviceToHost, cs));
  }
  ODEB200_CUDA(cudaEventRecord(g_arena.ev_done2, g_arena.stream2));
  ODEB200_CUDA(cudaStreamWaitEvent(st, g_arena.ev_done2, 0));
  if (n_chunks > 1) ODEB200_CUDA(cudaEventRecord(g_ev1, st));
  g_ev_pending = true;
  if (pts) {
    ODEB200_CUDA(cudaMemcpyAsync(io->pts_t, d.pts_t, (size_t)n * io->pts_cap * 8, cudaMemcpyDeviceToHost, st));
    ODEB200_CUDA(cudaMemcpyAsync(io->pts_y, d.pts_y, (size_t)n * io->pts_cap * D * 8, cudaMemcpyDeviceToHost, st));
    if (io->pts_n) ODEB200_CUDA(cudaMemcpyAsync(io->pts_n, d.pts_n, n * 4, cudaMemcpyDeviceToHost, st));
  }
  if (io->trace) {
    ODEB200_CUDA(cudaMemcpyAsync(io->trace, d.trace, (size_t)io->trace_cap * 32, cudaMemcpyDeviceToHost, st));
    if (io->trace_n) ODEB200_CUDA(cudaMemcpyAsync(io->trace_n, dtrace_n, 8, cudaMemcpyDeviceToHost, st));
  }
  ODEB200_CUDA(cudaStreamSynchronize(st));
  return 0;
}

// ---- single system with the reference's (tout, yout) contract ----------------
struct ode_b200_result {
  int dof = 0;
  int status = 0;
  long long nacc = 0, nrej = 0, nrhs = 0;
  std::vector<double> t, y;
};

int ode_b200_solve_single(const ode_b200_opts* opts, const double* y0, const double* params, const double* tspan,
                          int32_t n_tspan, ode_b200_result** out) {
  if (!opts || !y0 || !tspan || !out) return set_err(ODE_B200_E_INVALID, "null argument");
  *out = nullptr;
  if (opts->points == ODE_B200_POINTS_FINAL) return set_err(ODE_B200_E_POINTS, "Unrecognized option points==final");
  ode_b200_opts o = *opts;
  o.params_stride = 0;
  o.n_gpus = 1;  // one trajectory lives on one device
  long long cap = n_tspan + 1024;
  const bool fixed = is_fixed_step(o.method);
  if (fixed || o.points == ODE_B200_POINTS_SPECIFIED) cap = n_tspan;
  for (int tries = 0; tries < 8; tries++) {
    std::vector<double> pt(cap), py((size_t)cap * o.dof), yf(o.dof);
    double tf = 0;
    int32_t na = 0, nr = 0, nf = 0, st = 0, pn = 0;
    ode_b200_ensemble_io io;
    memset(&io, 0, sizeof(io));
    io.n_traj = 1;
    io.y0 = y0;
    io.params = params;
    io.tspan = tspan;
    io.n_tspan = n_tspan;
    io.t_final = &tf;
    io.y_final = yf.data();
    io.n_accept = &na;
    io.n_reject = &nr;
    io.n_rhs = &nf;
    io.status = &st;
    io.pts_t = pt.data();
    io.pts_y = py.data();
    io.pts_n = &pn;
    io.pts_cap = cap;
    io.trace_traj = -1;
    int rc = ode_b200_solve_ensemble(&o, &io);
    if (rc) return rc;
    if (st == ODE_B200_ST_OVERFLOW) {  // deterministic: rerun with the reported size
      cap = (long long)pn + 16;
      continue;
    }
    ode_b200_result* r = new (std::nothrow) ode_b200_result();
    if (!r) return set_err(ODE_B200_E_NOMEM, "out of memory");
    r->dof = o.dof;
    r->status = st;
    r->nacc = na;
    r->nrej = nr;
    r->nrhs = nf;
    r->t.assign(pt.begin(), pt.begin() + pn);
    r->y.assign(py.begin(), py.begin() + (size_t)pn * o.dof);
    *out = r;
    return 0;
  }
  return set_err(ODE_B200_E_NOMEM, "points buffer kept overflowing");
}
int64_t ode_b200_result_len(const ode_b200_result* r) { return r ? (int64_t)r->t.size() : 0; }
int32_t ode_b200_result_dof(const ode_b200_result* r) { return r ? r->dof : 0; }
int32_t ode_b200_result_status(const ode_b200_result* r) { return r ? r->status : ODE_B200_E_INVALID; }
int64_t ode_b200_result_counts(const ode_b200_result* r, int which) {
  if (!r) return 0;
  return which == 0 ? r->nacc : which == 1 ? r->nrej : r->nrhs;
}
int ode_b200_result_copy(const ode_b200_result* r, double* tout, double* yout) {
  if (!r) return set_err(ODE_B200_E_INVALID, "null result");
  if (tout) memcpy(tout, r->t.data(), r->t.size() * 8);
  if (yout) memcpy(yout, r->y.data(), r->y.size() * 8);
  return 0;
}
void ode_b200_result_free(ode_b200_result* r) { delete r; }

}  // extern "C"

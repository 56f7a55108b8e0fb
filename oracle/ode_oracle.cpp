// ode_oracle.cpp -- CPU restatement of ODE.jl's RK / Rosenbrock hot path.
//
// TEST INFRASTRUCTURE ONLY.  Nothing under ode.jl_b200/ may include, link or
// call this file; only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs use it, as the checker and the CPU
// baseline.  It follows /root/reference (SciML/ODE.jl v2.15.0) operation for
// operation; every function cites the lines it restates.
//
// PARITY PINNING.  The reference is pure Julia and Julia is not installed here
// (SURVEY.md F2), so this restatement is pinned against the reference's own
// known-answer tests (test/runtests.jl:40-70 analytic problems, :78-96 ROBER
// refsol at 2e-10, README.md:27) and cross-checked bit-for-bit against an
// independent pure-Python restatement (oracle/ode_oracle.py).  At the ulp level
// (Base `^`, `log10`, LAPACK getrf/getrs, BLAS nrm2 for dof >= 32) the
// reference's bits depend on the Julia version and are NOT pinned by any
// reference test: "parity unpinned" at that level.
//   mathmode 1: `^`/`log10` from the C library (closest stand-in for Julia's libm)
//   mathmode 0: the deterministic ode_b200_detmath.h versions (what the CUDA
//               kernels use; correctly rounded in all sampled cases)
//
// Build: g++ -O2 -ffp-contract=off -fno-fast-math -mfma -fopenmp (oracle/Makefile).
// -ffp-contract=off matters: the reference never fuses a*b+c.

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/ode_b200.h"
#include "../include/ode_b200_detmath.h"

namespace {

typedef std::vector<double> Vec;

// ---------------------------------------------------------------- Julia Base
// min/max propagate NaN (Base.min/max for Float64).
inline double jl_min(double a, double b) {
  if (a != a || b != b) return std::numeric_limits<double>::quiet_NaN();
  if (a < b) return a;
  if (b < a) return b;
  return std::signbit(a) ? a : b;
}
inline double jl_max(double a, double b) {
  if (a != a || b != b) return std::numeric_limits<double>::quiet_NaN();
  if (a > b) return a;
  if (b > a) return b;
  return std::signbit(a) ? b : a;
}
inline double jl_sign(double x) { return x > 0 ? 1.0 : (x < 0 ? -1.0 : x); }
// eps(x): spacing of floats at |x| (Base.eps(::Float64)).
inline double jl_eps(double x) {
  if (!std::isfinite(x)) return std::numeric_limits<double>::quiet_NaN();
  double ax = std::fabs(x);
  if (ax < 2.2250738585072014e-308) return 4.9406564584124654e-324;
  int e;
  std::frexp(ax, &e);  // ax = m * 2^e, m in [0.5, 1)  => exponent(x) = e - 1
  return std::ldexp(2.220446049250313e-16, e - 1);
}

struct MathCtx {
  int mode;  // 0 detmath, 1 libm
  double pow(double x, double y) const { return mode ? std::pow(x, y) : dm_pow(x, y); }
  double log10(double x) const { return mode ? std::log10(x) : dm_log10(x); }
};

// LinearAlgebra.norm(x, Inf): generic_normInf -- NaN-propagating max |x_i|.
inline double norm_inf(const double* x, int64_t n) {
  double m = std::fabs(x[0]);
  for (int64_t i = 1; i < n; i++) {
    double v = std::fabs(x[i]);
    m = ((m != m) || (m > v)) ? m : v;
  }
  return m;
}
// Error-free sum of squares for long vectors: double-double accumulation.
// (Julia calls BLAS nrm2 for n >= 32 whose bits are implementation defined;
// the CUDA large-system kernels use the same double-double scheme, so the
// rounded result is order independent in practice.)
inline double sumsq_dd(const double* x, int64_t n) {
  double hi = 0.0, lo = 0.0;
  for (int64_t i = 0; i < n; i++) {
    double p = x[i] * x[i];
    double pe = dm_fma(x[i], x[i], -p);
    dm_dd s = dm_two_sum(hi, p);
    lo = lo + (s.lo + pe);
    hi = s.hi;
  }
  return hi + lo;
}
// LinearAlgebra.norm(x) / norm(x, 2): generic_norm2 for n < 32 (sequential sum
// of squares, scaled variant when squares over/underflow).
inline double norm2(const double* x, int64_t n) {
  double maxabs = norm_inf(x, n);
  if (maxabs == 0.0 || std::isinf(maxabs)) return maxabs;
  if (n >= 32) {
    if (maxabs != maxabs) return maxabs;
    return std::sqrt(sumsq_dd(x, n));
  }
  double mm = maxabs * maxabs;
  if (std::isfinite((double)n * maxabs * maxabs) && mm != 0.0) {
    double sum = x[0] * x[0];
    for (int64_t i = 1; i < n; i++) sum += x[i] * x[i];
    return std::sqrt(sum);
  } else {
    double q = std::fabs(x[0]) / maxabs;
    double sum = q * q;
    for (int64_t i = 1; i < n; i++) {
      q = std::fabs(x[i]) / maxabs;
      sum += q * q;
    }
    return maxabs * std::sqrt(sum);
  }
}

// Elementwise loop, OpenMP-parallel only for long vectors (the large-system CPU baseline).  A plain
// branch, not an `if` clause: entering a (serialised) parallel region per small loop costs microseconds.
template <class Fn>
inline void par_for(int64_t n, Fn&& f) {
  if (n >= 65536) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++) f(i);
  } else {
    for (int64_t i = 0; i < n; i++) f(i);
  }
}

// ------------------------------------------------------------------ tableaus
// src/runge_kutta.jl:66-157, entries as reduced rationals; converted as
// Float64(num)/Float64(den) (Rational -> Float64, src/ODE.jl:163).
struct Q {
  long long n, d;
};
struct Tab {
  int S = 0, nb = 0;
  int order[2] = {0, 0};
  Vec a, b, c;  // a[S*S], b[nb*S], c[S]
  bool adaptive() const { return nb == 2; }  // :58
  bool fsal() const {                        // :62
    for (int j = 0; j < S; j++)
      if (a[(S - 1) * S + j] != b[j]) return false;
    return c[S - 1] == 1.0;
  }
  int min_order() const { return nb == 2 ? (order[0] < order[1] ? order[0] : order[1]) : order[0]; }
};
inline double qd(Q q) { return (double)q.n / (double)q.d; }

Tab make_tab(int method) {
  Tab t;
  auto fill = [&](int S, int nb, int o0, int o1, const Q* a, const Q* b, const Q* c) {
    t.S = S;
    t.nb = nb;
    t.order[0] = o0;
    t.order[1] = o1;
    t.a.resize(S * S);
    t.b.resize(nb * S);
    t.c.resize(S);
    for (int i = 0; i < S * S; i++) t.a[i] = qd(a[i]);
    for (int i = 0; i < nb * S; i++) t.b[i] = qd(b[i]);
    for (int i = 0; i < S; i++) t.c[i] = qd(c[i]);
  };
#define Z {0, 1}
  switch (method) {
    case ODE_B200_M_FEULER: {  // :66
      static const Q a[] = {Z}, b[] = {{1, 1}}, c[] = {Z};
      fill(1, 1, 1, 0, a, b, c);
    } break;
    case ODE_B200_M_MIDPOINT: {  // :71
      static const Q a[] = {Z, Z, {1, 2}, Z}, b[] = {Z, {1, 1}}, c[] = {Z, {1, 2}};
      fill(2, 1, 2, 0, a, b, c);
    } break;
    case ODE_B200_M_HEUN: {  // :77
      static const Q a[] = {Z, Z, {1, 1}, Z}, b[] = {{1, 2}, {1, 2}}, c[] = {Z, {1, 1}};
      fill(2, 1, 2, 0, a, b, c);
    } break;
    case ODE_B200_M_RK4: {  // :83
      static const Q a[] = {Z, Z, Z, Z, {1, 2}, Z, Z, Z, Z, {1, 2}, Z, Z, Z, Z, {1, 1}, Z};
      static const Q b[] = {{1, 6}, {1, 3}, {1, 3}, {1, 6}}, c[] = {Z, {1, 2}, {1, 2}, {1, 1}};
      fill(4, 1, 4, 0, a, b, c);
    } break;
    case ODE_B200_M_RK21: {  // :93
      static const Q a[] = {Z, Z, {1, 1}, Z}, b[] = {{1, 2}, {1, 2}, {1, 1}, Z}, c[] = {Z, {1, 1}};
      fill(2, 2, 2, 1, a, b, c);
    } break;
    case ODE_B200_M_RK23: {  // :101
      static const Q a[] = {Z, Z, Z, Z, {1, 2}, Z, Z, Z, Z, {3, 4}, Z, Z, {2, 9}, {1, 3}, {4, 9}, Z};
      static const Q b[] = {{7, 24}, {1, 4}, {1, 3}, {1, 8}, {2, 9}, {1, 3}, {4, 9}, Z};
      static const Q c[] = {Z, {1, 2}, {3, 4}, {1, 1}};
      fill(4, 2, 2, 3, a, b, c);
    } break;
    case ODE_B200_M_RK45_FE: {  // :112
      static const Q a[] = {Z, Z, Z, Z, Z, Z,
                            {1, 4}, Z, Z, Z, Z, Z,
                            {3, 32}, {9, 32}, Z, Z, Z, Z,
                            {1932, 2197}, {-7200, 2197}, {7296, 2197}, Z, Z, Z,
                            {439, 216}, {-8, 1}, {3680, 513}, {-845, 4104}, Z, Z,
                            {-8, 27}, {2, 1}, {-3544, 2565}, {1859, 4104}, {-11, 40}, Z};
      static const Q b[] = {{25, 216}, Z, {1408, 2565}, {2197, 4104}, {-1, 5}, Z,
                            {16, 135}, Z, {6656, 12825}, {28561, 56430}, {-9, 50}, {2, 55}};
      static const Q c[] = {Z, {1, 4}, {3, 8}, {12, 13}, {1, 1}, {1, 2}};
      fill(6, 2, 4, 5, a, b, c);
    } break;
    case ODE_B200_M_DOPRI5: {  // :124
      static const Q a[] = {Z, Z, Z, Z, Z, Z, Z,
                            {1, 5}, Z, Z, Z, Z, Z, Z,
                            {3, 40}, {9, 40}, Z, Z, Z, Z, Z,
                            {44, 45}, {-56, 15}, {32, 9}, Z, Z, Z, Z,
                            {19372, 6561}, {-25360, 2187}, {64448, 6561}, {-212, 729}, Z, Z, Z,
                            {9017, 3168}, {-355, 33}, {46732, 5247}, {49, 176}, {-5103, 18656}, Z, Z,
                            {35, 384}, Z, {500, 1113}, {125, 192}, {-2187, 6784}, {11, 84}, Z};
      static const Q b[] = {{35, 384}, Z, {500, 1113}, {125, 192}, {-2187, 6784}, {11, 84}, Z,
                            {5179, 57600}, Z, {7571, 16695}, {393, 640}, {-92097, 339200}, {187, 2100}, {1, 40}};
      static const Q c[] = {Z, {1, 5}, {3, 10}, {4, 5}, {8, 9}, {1, 1}, {1, 1}};
      fill(7, 2, 5, 4, a, b, c);
    } break;
    case ODE_B200_M_FEH78: {  // :140
      static const Q a[] = {
          Z, Z, Z, Z, Z, Z, Z, Z, Z, Z, Z, Z, Z,
          {2, 27}, Z, Z, Z, Z, Z, Z, Z, Z, Z, Z, Z, Z,
          {1, 36}, {1, 12}, Z, Z, Z, Z, Z, Z, Z, Z, Z, Z, Z,
          {1, 24}, Z, {1, 8}, Z, Z, Z, Z, Z, Z, Z, Z, Z, Z,
          {5, 12}, Z, {-25, 16}, {25, 16}, Z, Z, Z, Z, Z, Z, Z, Z, Z,
          {1, 20}, Z, Z, {1, 4}, {1, 5}, Z, Z, Z, Z, Z, Z, Z, Z,
          {-25, 108}, Z, Z, {125, 108}, {-65, 27}, {125, 54}, Z, Z, Z, Z, Z, Z, Z,
          {31, 300}, Z, Z, Z, {61, 225}, {-2, 9}, {13, 900}, Z, Z, Z, Z, Z, Z,
          {2, 1}, Z, Z, {-53, 6}, {704, 45}, {-107, 9}, {67, 90}, {3, 1}, Z, Z, Z, Z, Z,
          {-91, 108}, Z, Z, {23, 108}, {-976, 135}, {311, 54}, {-19, 60}, {17, 6}, {-1, 12}, Z, Z, Z, Z,
          {2383, 4100}, Z, Z, {-341, 164}, {4496, 1025}, {-301, 82}, {2133, 4100}, {45, 82}, {45, 164}, {18, 41}, Z, Z, Z,
          {3, 205}, Z, Z, Z, Z, {-6, 41}, {-3, 205}, {-3, 41}, {3, 41}, {6, 41}, Z, Z, Z,
          {-1777, 4100}, Z, Z, {-341, 164}, {4496, 1025}, {-289, 82}, {2193, 4100}, {51, 82}, {33, 164}, {12, 41}, Z, {1, 1}, Z};
      static const Q b[] = {{41, 840}, Z, Z, Z, Z, {34, 105}, {9, 35}, {9, 35}, {9, 280}, {9, 280}, {41, 840}, Z, Z,
                            Z, Z, Z, Z, Z, {34, 105}, {9, 35}, {9, 35}, {9, 280}, {9, 280}, Z, {41, 840}, {41, 840}};
      static const Q c[] = {Z, {2, 27}, {1, 9}, {1, 6}, {5, 12}, {1, 2}, {5, 6}, {1, 6}, {2, 3}, {1, 3}, {1, 1}, Z, {1, 1}};
      fill(13, 2, 7, 8, a, b, c);
    } break;
    case ODE_B200_M_CK45: {  // EXTENSION (not in the reference source; README.md:57 names it): Cash & Karp 1990, order (5,4)
      static const Q a[] = {Z, Z, Z, Z, Z, Z,
                            {1, 5}, Z, Z, Z, Z, Z,
                            {3, 40}, {9, 40}, Z, Z, Z, Z,
                            {3, 10}, {-9, 10}, {6, 5}, Z, Z, Z,
                            {-11, 54}, {5, 2}, {-70, 27}, {35, 27}, Z, Z,
                            {1631, 55296}, {175, 512}, {575, 13824}, {44275, 110592}, {253, 4096}, Z};
      static const Q b[] = {{37, 378}, Z, {250, 621}, {125, 594}, Z, {512, 1771},
                            {2825, 27648}, Z, {18575, 48384}, {13525, 55296}, {277, 14336}, {1, 4}};
      static const Q c[] = {Z, {1, 5}, {3, 10}, {3, 5}, {1, 1}, {7, 8}};
      fill(6, 2, 5, 4, a, b, c);
    } break;
    default:
      break;
  }
#undef Z
  return t;
}

// ---------------------------------------------------------- right-hand sides
// F(t, y) -> f.  Association is written out explicitly; the CUDA functors and
// the Python restatement state the same expressions independently.
typedef void (*UserRhsFn)(double t, const double* y, double* f, const double* p, int dof);
typedef void (*UserJacFn)(double t, const double* y, double* J, const double* p, int dof);
// periodic stencil of radius R: w[k] = y[i-R+k], parameters, stage time, grid index, ring length
typedef double (*UserPointFn)(const double* w, const double* p, double t, int64_t i, int64_t n);
struct Rhs {
  int id;
  int64_t dof;
  const double* p;
  UserRhsFn user = nullptr;  // run-time RHS (tests pass a ctypes callback evaluating the same expression)
  UserJacFn user_jac = nullptr;
  UserPointFn user_point = nullptr;  // run-time stencil of the large-system regime
  int user_radius = 1;
  mutable int64_t calls = 0;
  void operator()(double t, const double* y, double* f) const {
    calls++;
    if (user_point) {
      const int64_t n = dof;
      const int R = user_radius;
      double w[16];
      for (int64_t i = 0; i < n; i++) {
        for (int k = 0; k < 2 * R + 1; k++) w[k] = y[((i + k - R) % n + n) % n];
        f[i] = user_point(w, p, t, i, n);
      }
      return;
    }
    if (user) {
      user(t, y, f, p, (int)dof);
      return;
    }
    switch (id) {
      case ODE_B200_RHS_LINEAR:  // README.md:19  f(u,p,t) = 1.01*u
        f[0] = p[0] * y[0];
        break;
      case ODE_B200_RHS_CONST:  // test/runtests.jl:40
        f[0] = p[0];
        break;
      case ODE_B200_RHS_TIMELIN:  // test/runtests.jl:47  T(2)*t
        f[0] = p[0] * t;
        break;
      case ODE_B200_RHS_ROTATION:  // test/runtests.jl:67  [-y[2]; y[1]]
        f[0] = -y[1];
        f[1] = y[0];
        break;
      case ODE_B200_RHS_LORENZ: {  // examples/Lorenz_Attractor.ipynb cell 4
        double x = y[0], yy = y[1], z = y[2];
        f[0] = p[0] * (yy - x);
        f[1] = x * (p[1] - z) - yy;
        f[2] = x * yy - p[2] * z;
      } break;
      case ODE_B200_RHS_VDP: {
        double a = y[0], b = y[1];
        f[0] = b;
        f[1] = p[0] * ((1.0 - a * a) * b) - a;
      } break;
      case ODE_B200_RHS_ROBERTSON: {  // test/runtests.jl:81-87 (order: ydot1, ydot3, ydot2)
        double d1 = -p[0] * y[0] + p[2] * y[1] * y[2];
        double d3 = p[1] * y[1] * y[1];
        f[0] = d1;
        f[2] = d3;
        f[1] = -d1 - d3;
      } break;
      case ODE_B200_RHS_KEPLER: {
        double x = y[0], yy = y[1];
        double r2 = x * x + yy * yy;
        double r3 = r2 * std::sqrt(r2);
        f[0] = y[2];
        f[1] = y[3];
        f[2] = -x / r3;
        f[3] = -yy / r3;
      } break;
      case ODE_B200_RHS_REACTION_DIFFUSION: {
        const double kap = p[0], r = p[1];
        const int64_t n = dof;
        par_for(n, [&](int64_t i) {
          double um = y[i == 0 ? n - 1 : i - 1], u = y[i], up = y[i == n - 1 ? 0 : i + 1];
          f[i] = kap * ((um - 2.0 * u) + up) + (r * u) * (1.0 - u);
        });
      } break;
      default:
        break;
    }
  }
};

// Registered analytic Jacobians dF/dy (row-major n x n): what a user would pass as
// `jacobian = G(t,y)` (src/ODE.jl:262-263, :368-369).  Entry expressions are this
// repo's definition; the CUDA functors state the same expressions.
bool rhs_jacobian(const Rhs& F, double t, const double* y, double* J) {
  const double* p = F.p;
  if (F.user_jac) {
    F.user_jac(t, y, J, p, (int)F.dof);
    return true;
  }
  switch (F.id) {
    case ODE_B200_RHS_LINEAR: J[0] = p[0]; return true;
    case ODE_B200_RHS_CONST: J[0] = 0.0; return true;
    case ODE_B200_RHS_TIMELIN: J[0] = 0.0; return true;
    case ODE_B200_RHS_ROTATION: J[0] = 0.0; J[1] = -1.0; J[2] = 1.0; J[3] = 0.0; return true;
    case ODE_B200_RHS_LORENZ:
      J[0] = -p[0]; J[1] = p[0]; J[2] = 0.0;
      J[3] = p[1] - y[2]; J[4] = -1.0; J[5] = -y[0];
      J[6] = y[1]; J[7] = y[0]; J[8] = -p[2];
      return true;
    case ODE_B200_RHS_VDP:
      J[0] = 0.0; J[1] = 1.0;
      J[2] = p[0] * ((-2.0 * y[0]) * y[1]) - 1.0; J[3] = p[0] * (1.0 - y[0] * y[0]);
      return true;
    case ODE_B200_RHS_ROBERTSON: {
      const double a = p[2] * y[2], b = p[2] * y[1], c = (2.0 * p[1]) * y[1];
      J[0] = -p[0]; J[1] = a; J[2] = b;
      J[3] = p[0]; J[4] = -a - c; J[5] = -b;
      J[6] = 0.0; J[7] = c; J[8] = 0.0;
      return true;
    }
    case ODE_B200_RHS_KEPLER: {
      const double x = y[0], yy = y[1];
      const double r2 = x * x + yy * yy;
      const double r3 = r2 * std::sqrt(r2), r5 = r3 * r2;
      for (int i = 0; i < 16; i++) J[i] = 0.0;
      J[2] = 1.0; J[7] = 1.0;
      J[8] = (3.0 * x) * x / r5 - 1.0 / r3; J[9] = (3.0 * x) * yy / r5;
      J[12] = (3.0 * x) * yy / r5; J[13] = (3.0 * yy) * yy / r5 - 1.0 / r3;
      return true;
    }
    default: return false;
  }
}

// ------------------------------------------------------------- result record
struct Result {
  int dof = 0;
  std::vector<double> t;  // tout
  std::vector<double> y;  // yout, [len][dof]
  int64_t nacc = 0, nrej = 0, nrhs = 0;
  int status = ODE_B200_ST_OK;
  std::vector<double> trace;  // (t, dt, err, accepted) per attempt
  int err = 0;                // ODE_B200_E_*
};

struct Opt {
  double reltol, abstol, minstep, maxstep, initstep;
  int points;
  int64_t max_steps;
  bool want_trace;
  int jacobian = 0;  // 0: fdjacobian, 1: registered analytic Jacobian
};

// hinit, src/ODE.jl:85-112.  Returns signed h; sets tdir and f0.
int hinit(const Rhs& F, const MathCtx& M, const Vec& x0, double t0, double tend, int p, double reltol,
          double abstol, double* h_out, double* tdir_out, Vec& f0) {
  const int64_t n = (int64_t)x0.size();
  double tdir = jl_sign(tend - t0);                        // :87
  if (tdir == 0.0) return ODE_B200_E_ZERO_SPAN;            // :88
  double tau = jl_max(reltol * norm_inf(x0.data(), n), abstol);  // :89
  double d0 = norm_inf(x0.data(), n) / tau;                // :90
  f0.resize(n);
  F(t0, x0.data(), f0.data());                             // :91
  double d1 = norm_inf(f0.data(), n) / tau;                // :92
  double h0;
  if (d0 < 1e-5 || d1 < 1e-5)                              // :93
    h0 = 1e-6;
  else
    h0 = 0.01 * (d0 / d1);                                 // :96
  Vec x1(n), f1(n), df(n);
  double th = tdir * h0;
  for (int64_t i = 0; i < n; i++) x1[i] = x0[i] + th * f0[i];  // :100  x0 + tdir*h0*f0
  F(t0 + tdir * h0, x1.data(), f1.data());                 // :101
  for (int64_t i = 0; i < n; i++) df[i] = f1[i] - f0[i];
  double d2 = norm_inf(df.data(), n) / (tau * h0);         // :103
  double h1;
  double dm = jl_max(d1, d2);
  if (dm <= 1e-15) {                                       // :104
    h1 = jl_max(1e-6, 1e-3 * h0);                          // :105  T(10)^(-6), T(10)^(-3)*h0
  } else {
    double pw = -(2.0 + M.log10(dm)) / (double)(p + 1);    // :107
    h1 = M.pow(10.0, pw);                                  // :108
  }
  *h_out = tdir * jl_min(jl_min(100.0 * h0, h1), tdir * (tend - t0));  // :111
  *tdir_out = tdir;
  return 0;
}

// hermite_interp!, src/runge_kutta.jl:479-492.
inline void hermite(double* y, double tq, double t, double dt, const double* y0, const double* y1,
                    const double* f0, const double* f1, int64_t n) {
  double theta = (tq - t) / dt;  // :486
  for (int64_t i = 0; i < n; i++) {
    // :488-489
    y[i] = ((1.0 - theta) * y0[i] + theta * y1[i]) +
           (theta * (theta - 1.0)) *
               (((1.0 - 2.0 * theta) * (y1[i] - y0[i]) + ((theta - 1.0) * dt) * f0[i]) + (theta * dt) * f1[i]);
  }
}

// calc_next_k!, src/runge_kutta.jl:444-458.
inline void calc_next_k(std::vector<Vec>& ks, Vec& ytmp, const Vec& y, int s /*0-based*/, const Rhs& F, double t,
                        double dt, const Tab& bt) {
  const int64_t n = (int64_t)y.size();
  const int S = bt.S;
  // (loop order swapped to d-outer for long vectors: per component the additions stay in ss order, :453-455)
  par_for(n, [&](int64_t d) {
    double v = y[d];                                                        // :452
    for (int ss = 0; ss < s; ss++) v += dt * ks[ss][d] * bt.a[s * S + ss];  // :454 (dt*k)*a
    ytmp[d] = v;
  });
  ks[s].resize(n);
  F(t + bt.c[s] * dt, ytmp.data(), ks[s].data());  // :456
}

// oderk_adapt, src/runge_kutta.jl:232-369 (with rk_embedded_step! :371-399 and
// stepsize_hw92! :401-442 inlined where they are called).
void oderk_adapt(const Rhs& F, const MathCtx& M, const Tab& bt, const Vec& y0, const Vec& tspan_in, const Opt& o,
                 Result& R) {
  const int S = bt.S;
  const int64_t dof = (int64_t)y0.size();
  R.dof = (int)dof;
  if (!bt.adaptive()) {  // :249
    R.err = ODE_B200_E_NOT_ADAPTIVE;
    return;
  }
  const int order = bt.min_order();  // :252
  const int timeout_const = 5;       // :253
  const Vec& tsf = tspan_in;         // tspan_fixed
  const int64_t nfixed = (int64_t)tsf.size();  // :277
  double t = tsf[0], tstart = tsf[0], tend = tsf[nfixed - 1];  // :259-261
  Vec y(y0), ytrial(dof), yerr(dof), ytmp(dof);
  std::vector<Vec> ks(S);
  if (o.points != ODE_B200_POINTS_ALL && o.points != ODE_B200_POINTS_SPECIFIED && o.points != ODE_B200_POINTS_FINAL) {
    R.err = ODE_B200_E_POINTS;  // :289
    return;
  }
  // outputs: ys[1] = y0 (:280); tspan = [tstart] for :all (:285)
  std::vector<Vec> ys;
  Vec tout;
  if (o.points == ODE_B200_POINTS_SPECIFIED) {
    ys.assign(nfixed, Vec(dof, 0.0));
    ys[0] = y0;
    tout = tsf;
  } else {
    ys.push_back(y0);
    tout.push_back(tstart);
  }
  int64_t iter_fixed = 2;  // :286 (1-based index into tspan_fixed)
  double dt, tdir;
  int e = hinit(F, M, y, tstart, tend, order, o.reltol, o.abstol, &dt, &tdir, ks[0]);  // :292
  if (e) {
    R.err = e;
    return;
  }
  if (o.initstep != 0) {  // :293-295
    if (jl_sign(o.initstep) == tdir)
      dt = o.initstep;
    else {
      R.err = ODE_B200_E_INITSTEP;
      return;
    }
  }
  bool islaststep = std::fabs(t + dt - tend) <= jl_eps(tend);  // :302
  int timeout = 0;                                             // :303
  int64_t iter = 2;                                            // :304 (1-based)
  int64_t attempts = 0;
  const double fac = 0.8, facmin = 1.0 / 5.0;  // :422-424
  while (true) {                                // :305
    if (o.max_steps > 0 && attempts >= o.max_steps) {
      R.status = ODE_B200_ST_MAXSTEPS;
      break;
    }
    attempts++;
    // ---- rk_embedded_step! :371-399
    par_for(dof, [&](int64_t d) {  // :382-387
      ytrial[d] = 0.0;
      yerr[d] = 0.0;
      ytrial[d] += bt.b[0 * S + 0] * ks[0][d];
      yerr[d] += bt.b[1 * S + 0] * ks[0][d];
    });
    for (int s = 1; s < S; s++) {  // :388-394
      calc_next_k(ks, ytmp, y, s, F, t, dt, bt);
      par_for(dof, [&](int64_t d) {
        ytrial[d] += bt.b[0 * S + s] * ks[s][d];
        yerr[d] += bt.b[1 * S + s] * ks[s][d];
      });
    }
    par_for(dof, [&](int64_t d) {  // :395-398
      yerr[d] = dt * (ytrial[d] - yerr[d]);
      ytrial[d] = y[d] + dt * ytrial[d];
    });
    // ---- stepsize_hw92! :401-442
    double err, newdt;
    bool nanout = false;
    for (int64_t d = 0; d < dof; d++) {  // :430-434
      if (ytrial[d] != ytrial[d]) {      // isoutofdomain = isnan, src/ODE.jl:116
        nanout = true;
        break;
      }
      yerr[d] = yerr[d] / (o.abstol + jl_max(std::fabs(y[d]), std::fabs(ytrial[d])) * o.reltol);  // :433
    }
    if (nanout) {
      err = 10.0;  // :432
      newdt = dt * facmin;
      timeout = 5;
    } else {
      err = norm2(yerr.data(), dof);  // :435
      double nd = jl_min(o.maxstep, (tdir * dt) * jl_max(facmin, fac * M.pow(1.0 / err, 1.0 / (double)(order + 1))));  // :436
      if (timeout > 0) {  // :437-440
        nd = jl_min(nd, dt);
        timeout -= 1;
      }
      newdt = tdir * nd;  // :441
    }
    if (o.want_trace) {
      R.trace.push_back(t);
      R.trace.push_back(dt);
      R.trace.push_back(err);
      R.trace.push_back(err <= 1 ? 1.0 : 0.0);
    }
    if (err <= 1) {  // :312 accept
      R.nacc++;
      Vec f1;
      const Vec& f0 = ks[0];
      if (bt.fsal())  // :320
        f1 = ks[S - 1];
      else {
        f1.resize(dof);
        F(t + dt, ytrial.data(), f1.data());
      }
      if (o.points == ODE_B200_POINTS_SPECIFIED) {  // :321-326
        while (iter - 1 < nfixed && (tdir * tsf[iter - 1] < tdir * (t + dt) || islaststep)) {
          hermite(ys[iter - 1].data(), tsf[iter - 1], t, dt, y.data(), ytrial.data(), f0.data(), f1.data(), dof);
          iter++;
        }
      } else if (o.points == ODE_B200_POINTS_ALL) {  // :327-340
        while (iter_fixed - 1 < nfixed && tdir * t < tdir * tsf[iter_fixed - 1] &&
               tdir * tsf[iter_fixed - 1] < tdir * (t + dt)) {
          Vec yo(dof);
          hermite(yo.data(), tsf[iter_fixed - 1], t, dt, y.data(), ytrial.data(), f0.data(), f1.data(), dof);
          ys.push_back(yo);
          tout.push_back(tsf[iter_fixed - 1]);
          iter_fixed++;
          iter++;
        }
        ys.push_back(ytrial);
        tout.push_back(t + dt);
        iter++;
      } else {  // FINAL (extension): keep only the latest accepted state
        ys.resize(1);
        tout.resize(1);
        ys[0] = ytrial;
        tout[0] = t + dt;
      }
      ks[0] = f1;             // :341
      if (islaststep) break;  // :344
      std::swap(y, ytrial);   // :347
      t += dt;                // :350
      dt = newdt;             // :351
      if (tdir * (t + dt + dt / 100) >= tdir * tend) {  // :354
        dt = tend - t;
        islaststep = true;
      }
    } else if (std::fabs(newdt) < o.minstep) {  // :358
      R.status = ODE_B200_ST_MINSTEP;           // "Warning: dt < minstep.  Stopping."
      break;
    } else {  // :361-366 reject
      islaststep = false;
      R.nrej++;
      dt = newdt;
      timeout = timeout_const;
      if (!std::isfinite(dt)) {  // guard (the reference loops forever)
        R.status = ODE_B200_ST_NONFINITE;
        break;
      }
    }
  }
  R.nrhs = F.calls;
  R.t = tout;
  R.y.resize(ys.size() * dof);
  for (size_t i = 0; i < ys.size(); i++)
    for (int64_t d = 0; d < dof; d++) R.y[i * dof + d] = ys[i][d];
}

// oderk_fixed, src/runge_kutta.jl:176-212.
void oderk_fixed(const Rhs& F, const Tab& bt, const Vec& y0, const Vec& tspan, Result& R) {
  const int S = bt.S;
  const int64_t dof = (int64_t)y0.size();
  const int64_t nt = (int64_t)tspan.size();
  R.dof = (int)dof;
  std::vector<Vec> ys(nt, Vec(dof, 0.0));
  ys[0] = y0;  // :193
  std::vector<Vec> ks(S);
  Vec ytmp(dof);
  for (int64_t i = 0; i + 1 < nt; i++) {  // :201
    double dt = tspan[i + 1] - tspan[i];  // :202
    ys[i + 1] = ys[i];                    // :203
    for (int s = 0; s < S; s++) {         // :204
      calc_next_k(ks, ytmp, ys[i], s, F, tspan[i], dt, bt);
      for (int64_t d = 0; d < dof; d++) ys[i + 1][d] += dt * bt.b[s] * ks[s][d];  // :207
    }
  }
  R.nacc = nt - 1;
  R.nrhs = F.calls;
  R.t = tspan;
  R.y.resize(nt * dof);
  for (int64_t i = 0; i < nt; i++)
    for (int64_t d = 0; d < dof; d++) R.y[i * dof + d] = ys[i][d];
}

// fdjacobian, src/ODE.jl:232-243.  J is dof x dof row-major.
void fdjacobian(const Rhs& F, const Vec& x, double t, Vec& J) {
  const int64_t n = (int64_t)x.size();
  Vec ftx(n), xp(n), fp(n), dx(n);
  F(t, x.data(), ftx.data());  // :233
  J.assign(n * n, 0.0);
  for (int64_t j = 0; j < n; j++) {  // :236
    for (int64_t i = 0; i < n; i++) dx[i] = 0.0;
    dx[j] = (x[j] + (x[j] == 0 ? 1.0 : 0.0)) / 100;  // :239
    for (int64_t i = 0; i < n; i++) xp[i] = x[i] + dx[i];
    F(t, xp.data(), fp.data());
    for (int64_t i = 0; i < n; i++) J[i * n + j] = (fp[i] - ftx[i]) / dx[j];  // :240
  }
}

// lu(A) as LAPACK dgetrf does it for small matrices (unblocked right-looking,
// partial pivoting on the first max |.|, scaling by the reciprocal pivot), and
// getrs (row swaps, unit-lower forward substitution, upper back substitution
// with division), reference-BLAS operation order.  Julia: LinearAlgebra.lu,
// `\` -> LAPACK.getrs!.  The exact bits of OpenBLAS are not pinned (header).
struct LU {
  int64_t n;
  Vec a;
  std::vector<int> piv;
  bool singular = false;
};
void lu_factor(LU& f) {
  const int64_t n = f.n;
  Vec& a = f.a;
  f.piv.resize(n);
  for (int64_t j = 0; j < n; j++) {
    int64_t p = j;
    double mx = std::fabs(a[j * n + j]);
    for (int64_t i = j + 1; i < n; i++) {
      double v = std::fabs(a[i * n + j]);
      if (v > mx) {
        mx = v;
        p = i;
      }
    }
    f.piv[j] = (int)p;
    if (a[p * n + j] == 0.0) {
      f.singular = true;
      continue;
    }
    if (p != j)
      for (int64_t c = 0; c < n; c++) std::swap(a[j * n + c], a[p * n + c]);
    double rp = 1.0 / a[j * n + j];
    for (int64_t i = j + 1; i < n; i++) a[i * n + j] = a[i * n + j] * rp;
    for (int64_t c = j + 1; c < n; c++) {
      double u = a[j * n + c];
      for (int64_t i = j + 1; i < n; i++) a[i * n + c] = a[i * n + c] - a[i * n + j] * u;
    }
  }
}
void lu_solve(const LU& f, Vec& b) {
  const int64_t n = f.n;
  const Vec& a = f.a;
  for (int64_t j = 0; j < n; j++)
    if (f.piv[j] != j) std::swap(b[j], b[f.piv[j]]);
  for (int64_t j = 0; j < n; j++)  // L, unit diagonal (column sweep)
    for (int64_t i = j + 1; i < n; i++) b[i] = b[i] - b[j] * a[i * n + j];
  for (int64_t j = n - 1; j >= 0; j--) {  // U
    b[j] = b[j] / a[j * n + j];
    for (int64_t i = 0; i < j; i++) b[i] = b[i] - b[j] * a[i * n + j];
  }
}

// ode23s, src/ODE.jl:252-361.
void ode23s(const Rhs& F, const MathCtx& M, const Vec& y0, const Vec& tspan, const Opt& o, Result& R) {
  const int64_t n = (int64_t)y0.size();
  R.dof = (int)n;
  const double d = 1.0 / (2.0 + std::sqrt(2.0));  // :270
  const double e32 = 6.0 + std::sqrt(2.0);        // :271
  double t = tspan[0];                            // :275
  const double tfinal = tspan[tspan.size() - 1];  // :277
  double h = o.initstep, tdir;                    // :279
  Vec F0(n);
  if (h == 0.0) {  // :280-282
    int e = hinit(F, M, y0, t, tfinal, 3, o.reltol, o.abstol, &h, &tdir, F0);
    if (e) {
      R.err = e;
      return;
    }
  } else {  // :284-285
    tdir = jl_sign(tfinal - t);
    F(t, y0.data(), F0.data());
  }
  h = tdir * jl_min(std::fabs(h), o.maxstep);  // :287
  Vec y(y0);
  Vec tout;
  std::vector<Vec> yout;
  tout.push_back(t);  // :290-291
  yout.push_back(y);
  Vec J;
  auto jac = [&](const Vec& yy, double tt) {  // :262-267
    if (o.jacobian) {
      J.assign(n * n, 0.0);
      rhs_jacobian(F, tt, yy.data(), J.data());
    } else {
      fdjacobian(F, yy, tt, J);
    }
  };
  jac(y, t);  // :294
  Vec Fa(n), T(n), k1(n), k2(n), k3(n), F1(n), F2(n), ynew(n), tmp(n);
  int64_t attempts = 0;
  while (std::fabs(t - tfinal) > 0 && o.minstep < std::fabs(h)) {  // :296
    if (o.max_steps > 0 && attempts >= o.max_steps) {
      R.status = ODE_B200_ST_MAXSTEPS;
      break;
    }
    attempts++;
    if (std::fabs(t - tfinal) < std::fabs(h)) h = tfinal - t;  // :297-299
    LU W;  // :301-310  W = lu(I - h*d*J)
    W.n = n;
    W.a.resize(n * n);
    double hd = h * d;
    for (int64_t i = 0; i < n; i++)
      for (int64_t j = 0; j < n; j++) {
        double m = -(hd * J[i * n + j]);
        W.a[i * n + j] = (i == j) ? (m + 1.0) : m;
      }
    lu_factor(W);
    if (W.singular) {
      R.status = ODE_B200_ST_SINGULAR;
      break;
    }
    // :313  T = h*d*(F(t + h/100, y) - F0)/(h/100)
    F(t + h / 100, y.data(), Fa.data());
    for (int64_t i = 0; i < n; i++) T[i] = (hd * (Fa[i] - F0[i])) / (h / 100);
    for (int64_t i = 0; i < n; i++) k1[i] = F0[i] + T[i];  // :316
    lu_solve(W, k1);
    double hh = 0.5 * h;
    for (int64_t i = 0; i < n; i++) tmp[i] = y[i] + hh * k1[i];  // :317
    F(t + 0.5 * h, tmp.data(), F1.data());
    for (int64_t i = 0; i < n; i++) k2[i] = F1[i] - k1[i];  // :318
    lu_solve(W, k2);
    for (int64_t i = 0; i < n; i++) k2[i] = k2[i] + k1[i];
    for (int64_t i = 0; i < n; i++) ynew[i] = y[i] + h * k2[i];  // :319
    F(t + h, ynew.data(), F2.data());                            // :320
    for (int64_t i = 0; i < n; i++)                              // :321
      k3[i] = ((F2[i] - e32 * (k2[i] - F1[i])) - 2 * (k1[i] - F0[i])) + T[i];
    lu_solve(W, k3);
    for (int64_t i = 0; i < n; i++) tmp[i] = (k1[i] - 2 * k2[i]) + k3[i];
    double err = (std::fabs(h) / 6) * norm2(tmp.data(), n);  // :323
    double delta = jl_max(o.reltol * jl_max(norm2(y.data(), n), norm2(ynew.data(), n)), o.abstol);  // :324
    bool acc = err <= delta;
    if (o.want_trace) {
      R.trace.push_back(t);
      R.trace.push_back(h);
      R.trace.push_back(err / delta);
      R.trace.push_back(acc ? 1.0 : 0.0);
    }
    if (acc) {  // :327
      R.nacc++;
      if (o.points == ODE_B200_POINTS_FINAL) {
        tout.resize(1);
        yout.resize(1);
        tout[0] = t + h;
        yout[0] = ynew;
      } else {
        // :329-340  tspan points in (t, t+h]
        for (size_t q = 0; q < tspan.size(); q++) {
          double toi = tspan[q];
          if (toi > t && toi <= t + h) {
            double s = (toi - t) / h;  // :334
            Vec yo(n);
            double den = 1 - 2 * d;
            for (int64_t i = 0; i < n; i++)  // :338
              yo[i] = y[i] + h * (((k1[i] * s) * (1 - s)) / den + ((k2[i] * s) * (s - 2 * d)) / den);
            tout.push_back(toi);
            yout.push_back(yo);
          }
        }
        if (o.points == ODE_B200_POINTS_ALL && tout.back() != t + h) {  // :341-345
          tout.push_back(t + h);
          yout.push_back(ynew);
        }
      }
      t = t + h;  // :348-352
      y = ynew;
      F0 = F2;
      jac(y, t);
    } else {
      R.nrej++;
    }
    h = tdir * jl_min(o.maxstep, (std::fabs(h) * 0.8) * M.pow(delta / err, 1.0 / 3));  // :357
  }
  // the reference returns silently when the loop guard fails short of tfinal; report why
  if (R.status == ODE_B200_ST_OK && std::fabs(t - tfinal) > 0)
    R.status = std::isfinite(h) ? ODE_B200_ST_MINSTEP : ODE_B200_ST_NONFINITE;
  R.nrhs = F.calls;
  R.t = tout;
  R.y.resize(yout.size() * n);
  for (size_t i = 0; i < yout.size(); i++)
    for (int64_t k = 0; k < n; k++) R.y[i * n + k] = yout[i][k];
}

// oderosenbrock, src/ODE.jl:366-404, coefficients :408-431.
struct RosCoef {
  double gamma, a[4][4], b[4], c[4][4];
};
const RosCoef& ros_coef(int method) {
  static const RosCoef kr = {0.231,
                             {{0, 0, 0, 0}, {2, 0, 0, 0}, {4.452470820736, 4.16352878860, 0, 0}, {4.452470820736, 4.16352878860, 0, 0}},
                             {3.95750374663, 4.62489238836, 0.617477263873, 1.28261294568},
                             {{0, 0, 0, 0}, {-5.07167533877, 0, 0, 0}, {6.02015272865, 0.1597500684673, 0, 0},
                              {-1.856343618677, -8.50538085819, -2.08407513602, 0}}};
  static const RosCoef sh = {0.5,
                             {{0, 0, 0, 0}, {2, 0, 0, 0}, {48.0 / 25, 6.0 / 25, 0, 0}, {48.0 / 25, 6.0 / 25, 0, 0}},
                             {19.0 / 9, 1.0 / 2, 25.0 / 108, 125.0 / 108},
                             {{0, 0, 0, 0}, {-8, 0, 0, 0}, {372.0 / 25, 12.0 / 5, 0, 0}, {-112.0 / 125, -54.0 / 125, -2.0 / 5, 0}}};
  return method == ODE_B200_M_ODE4S_KR ? kr : sh;
}
void oderosenbrock(const Rhs& F, int method, const Vec& x0, const Vec& tspan, const Opt& o, Result& R) {
  const RosCoef& k = ros_coef(method);
  const int64_t n = (int64_t)x0.size(), nt = (int64_t)tspan.size();
  R.dof = (int)n;
  std::vector<Vec> x(nt, Vec(n, 0.0));
  x[0] = x0;
  Vec J, f(n), arg(n), dx(n), dF(n);
  Vec g[4];
  for (int64_t st = 0; st + 1 < nt; st++) {  // :379
    const double ts = tspan[st], hs = tspan[st + 1] - tspan[st];
    const Vec& xs = x[st];
    if (o.jacobian) {  // :383
      J.assign(n * n, 0.0);
      rhs_jacobian(F, ts, xs.data(), J.data());
    } else {
      fdjacobian(F, xs, ts, J);
    }
    LU W;  // :385  jac = I/(gamma*hs) - dFdx   (every `jac \ v` re-factorises the same matrix)
    W.n = n;
    W.a.resize(n * n);
    const double lam = 1.0 / (k.gamma * hs);
    for (int64_t i = 0; i < n; i++)
      for (int64_t j = 0; j < n; j++) W.a[i * n + j] = (i == j) ? (-J[i * n + j] + lam) : -J[i * n + j];
    lu_factor(W);
    if (W.singular) {  // Julia raises SingularException; report the truncated solution instead
      R.status = ODE_B200_ST_SINGULAR;
      R.t.assign(tspan.begin(), tspan.begin() + st + 1);
      R.y.resize((st + 1) * n);
      for (int64_t i = 0; i <= st; i++)
        for (int64_t d = 0; d < n; d++) R.y[i * n + d] = x[i][d];
      R.nrhs = F.calls;
      return;
    }
    F(ts + k.b[0] * hs, xs.data(), f.data());  // :388
    g[0] = f;
    lu_solve(W, g[0]);
    for (int64_t d = 0; d < n; d++) x[st + 1][d] = xs[d] + k.b[0] * g[0][d];  // :389
    for (int i = 1; i < 4; i++) {  // :391
      for (int64_t d = 0; d < n; d++) {
        dx[d] = 0.0;
        dF[d] = 0.0;
      }
      for (int j = 0; j < i; j++)  // :394-397
        for (int64_t d = 0; d < n; d++) {
          dx[d] += k.a[i][j] * g[j][d];
          dF[d] += k.c[i][j] * g[j][d];
        }
      for (int64_t d = 0; d < n; d++) arg[d] = xs[d] + dx[d];
      F(ts + k.b[i] * hs, arg.data(), f.data());  // :398
      g[i].resize(n);
      for (int64_t d = 0; d < n; d++) g[i][d] = f[d] + dF[d] / hs;
      lu_solve(W, g[i]);
      for (int64_t d = 0; d < n; d++) x[st + 1][d] += k.b[i] * g[i][d];  // :399
    }
    R.nacc++;
  }
  R.nrhs = F.calls;
  R.t = tspan;
  R.y.resize(nt * n);
  for (int64_t i = 0; i < nt; i++)
    for (int64_t d = 0; d < n; d++) R.y[i * n + d] = x[i][d];
}

// ode_ms (src/ODE.jl:175-213): ode4ms = order 4, ode5ms = order 5, orders 1-3 through ODE.ode_ms.
// b for order <= 4 is ms_coefficients4 (:440-443); for order 5 the reference fills b[5, :] in floating point (:183-194):
//   p_int = polyint(poly(diagm(0 => -[0:j-1; j+1:s-1])));  b[s, s-j] = (-1)^j / factorial(j) / factorial(s-1-j) * polyval(p_int, 1)
// restated with poly = product of (x - r) over the roots in the order listed (small integers: exact), polyint = a_k/(k+1),
// polyval = Horner from the leading coefficient, the prefactor evaluated left to right.  Polynomials.jl is not vendored
// and Julia is absent: the bits of this row are UNPINNED against the reference (exact values [251,-1274,2616,-2774,1901]/720).
int ms_order_of(int method) {
  switch (method) {
    case ODE_B200_M_ODE_MS1: return 1;
    case ODE_B200_M_ODE_MS2: return 2;
    case ODE_B200_M_ODE_MS3: return 3;
    case ODE_B200_M_ODE4MS: return 4;
    case ODE_B200_M_ODE5MS: return 5;
    default: return 0;
  }
}
void ms_coefficients(int order, double (&b)[5][5]) {
  static const double b4[4][4] = {{1, 0, 0, 0},
                                  {-1.0 / 2, 3.0 / 2, 0, 0},
                                  {5.0 / 12, -4.0 / 3, 23.0 / 12, 0},
                                  {-9.0 / 24, 37.0 / 24, -59.0 / 24, 55.0 / 24}};
  for (int i = 0; i < 5; i++)
    for (int j = 0; j < 5; j++) b[i][j] = (i < 4 && j < 4) ? b4[i][j] : 0.0;
  for (int s = 5; s <= order; s++) {
    for (int j = 0; j <= s - 1; j++) {
      std::vector<double> pc(1, 1.0);  // ascending coefficients
      for (int r = 0; r <= s - 1; r++) {
        if (r == j) continue;
        const double root = -(double)r;
        std::vector<double> q(pc.size() + 1, 0.0);
        for (size_t k = 0; k < pc.size(); k++) {
          q[k + 1] += pc[k];
          q[k] += -root * pc[k];
        }
        pc = q;
      }
      std::vector<double> pi(pc.size() + 1, 0.0);
      for (size_t k = 0; k < pc.size(); k++) pi[k + 1] = pc[k] / (double)(k + 1);
      double v = pi.back();
      for (int k = (int)pi.size() - 2; k >= 0; k--) v = pi[k] + 1.0 * v;
      double fj = 1.0, fs = 1.0;
      for (int k = 2; k <= j; k++) fj *= k;
      for (int k = 2; k <= s - 1 - j; k++) fs *= k;
      b[s - 1][s - j - 1] = ((j % 2) ? -1.0 : 1.0) / fj / fs * v;
    }
  }
}
void ode_ms(const Rhs& F, const Vec& x0, const Vec& tspan, int order, Result& R) {
  double b[5][5];
  ms_coefficients(order, b);
  const int64_t n = (int64_t)x0.size(), nt = (int64_t)tspan.size();
  R.dof = (int)n;
  std::vector<Vec> x(nt, Vec(n, 0.0)), xdot(nt, Vec(n, 0.0));
  x[0] = x0;
  for (int64_t i = 0; i + 1 < nt; i++) {  // :198 (0-based i; the reference's i is i+1)
    const double h = tspan[i + 1] - tspan[i];
    const int so = (int)std::min<int64_t>(i + 1, order);  // :200
    F(tspan[i], x[i].data(), xdot[i].data());            // :201
    x[i + 1] = x[i];                                      // :203
    for (int j = 0; j < so; j++)                          // :204-206
      for (int64_t d = 0; d < n; d++) x[i + 1][d] += h * b[so - 1][j] * xdot[i - (so - 1) + j][d];
  }
  R.nacc = nt - 1;
  R.nrhs = F.calls;
  R.t = tspan;
  R.y.resize(nt * n);
  for (int64_t i = 0; i < nt; i++)
    for (int64_t d = 0; d < n; d++) R.y[i * n + d] = x[i][d];
}

thread_local UserRhsFn g_user_rhs = nullptr;
thread_local UserJacFn g_user_jac = nullptr;
thread_local UserPointFn g_user_point = nullptr;
thread_local int g_user_radius = 1;

void solve_one(const ode_b200_opts* op, const double* y0, const double* params, const double* tspan, int n_tspan,
               bool want_trace, Result& R) {
  Rhs F{op->rhs, op->dof, params};
  if (op->rhs >= ODE_B200_STENCIL_USER_BASE) {
    F.user_point = g_user_point;
    F.user_radius = g_user_radius;
  } else if (op->rhs >= ODE_B200_RHS_USER_BASE) {
    F.user = g_user_rhs;
    F.user_jac = g_user_jac;
  }
  MathCtx M{op->mathmode};
  Opt o{op->reltol, op->abstol, op->minstep, op->maxstep, op->initstep, op->points,
        op->max_steps > 0 ? op->max_steps : 100000000, want_trace, op->jacobian};
  Vec y(y0, y0 + op->dof), ts(tspan, tspan + n_tspan);
  if (op->method == ODE_B200_M_ODE23S) {
    ode23s(F, M, y, ts, o, R);
  } else if (op->method == ODE_B200_M_ODE4S_S || op->method == ODE_B200_M_ODE4S_KR) {
    oderosenbrock(F, op->method, y, ts, o, R);
  } else if (ms_order_of(op->method) > 0) {
    ode_ms(F, y, ts, ms_order_of(op->method), R);
  } else {
    Tab bt = make_tab(op->method);
    if (bt.S == 0) {
      R.err = ODE_B200_E_INVALID;
      return;
    }
    if (bt.adaptive())
      oderk_adapt(F, M, bt, y, ts, o, R);
    else
      oderk_fixed(F, bt, y, ts, R);
  }
}

}  // namespace

// ------------------------------------------------------------------ C exports
extern "C" {

struct oracle_result {
  Result r;
};

double oracle_ms_coefficient(int order, int j) {  // 1-based like the reference's b[order, j]
  double b[5][5];
  if (order < 1 || order > 5 || j < 1 || j > order) return NAN;
  ms_coefficients(order, b);
  return b[order - 1][j - 1];
}
// run-time RHS for opts.rhs >= ODE_B200_RHS_USER_BASE (single-trajectory solves of the calling thread)
void oracle_set_user_rhs(UserRhsFn f, UserJacFn j) {
  g_user_rhs = f;
  g_user_jac = j;
}
void oracle_set_user_stencil(UserPointFn f, int radius) {
  g_user_point = f;
  g_user_radius = radius >= 1 && radius <= 7 ? radius : 1;
}

int oracle_solve(const ode_b200_opts* op, const double* y0, const double* params, const double* tspan, int n_tspan,
                 int want_trace, oracle_result** out) {
  oracle_result* h = new oracle_result();
  solve_one(op, y0, params, tspan, n_tspan, want_trace != 0, h->r);
  *out = h;
  return h->r.err;
}
int64_t oracle_result_len(const oracle_result* h) { return (int64_t)h->r.t.size(); }
int64_t oracle_result_trace_len(const oracle_result* h) { return (int64_t)h->r.trace.size() / 4; }
int oracle_result_status(const oracle_result* h) { return h->r.status; }
int64_t oracle_result_count(const oracle_result* h, int which) {
  return which == 0 ? h->r.nacc : which == 1 ? h->r.nrej : h->r.nrhs;
}
void oracle_result_copy(const oracle_result* h, double* t, double* y, double* trace) {
  if (t) memcpy(t, h->r.t.data(), h->r.t.size() * 8);
  if (y) memcpy(y, h->r.y.data(), h->r.y.size() * 8);
  if (trace) memcpy(trace, h->r.trace.data(), h->r.trace.size() * 8);
}
void oracle_result_free(oracle_result* h) { delete h; }

// Ensemble: final state + counters per trajectory, OpenMP over trajectories.
// This is also the CPU baseline (bench.py cpu_baseline / --impl reference).
int oracle_solve_ensemble(const ode_b200_opts* op, int64_t n_traj, const double* y0, const double* params,
                          const double* tspan, int n_tspan, double* t_final, double* y_final, int32_t* nacc,
                          int32_t* nrej, int32_t* nrhs, int32_t* status, int n_threads) {
  int err = 0;
  const int dof = op->dof;
#pragma omp parallel for schedule(dynamic, 16) num_threads(n_threads > 0 ? n_threads : 1)
  for (int64_t i = 0; i < n_traj; i++) {
    Result R;
    ode_b200_opts o = *op;
    o.points = ODE_B200_POINTS_FINAL;
    const double* p = params + (op->params_stride ? i * (int64_t)op->params_stride : 0);
    solve_one(&o, y0 + i * dof, p, tspan, n_tspan, false, R);
    if (R.err) {
#pragma omp critical
      err = R.err;
      continue;
    }
    size_t last = R.t.size() - 1;
    if (t_final) t_final[i] = R.t[last];
    if (y_final)
      for (int d = 0; d < dof; d++) y_final[i * dof + d] = R.y[last * dof + d];
    if (nacc) nacc[i] = (int32_t)R.nacc;
    if (nrej) nrej[i] = (int32_t)R.nrej;
    if (nrhs) nrhs[i] = (int32_t)R.nrhs;
    if (status) status[i] = R.status;
  }
  return err;
}

double oracle_tableau(int method, int which, int i, int j) {
  Tab t = make_tab(method);
  if (t.S == 0) return std::numeric_limits<double>::quiet_NaN();
  if (which == 0) return t.a[i * t.S + j];
  if (which == 1) return t.b[i * t.S + j];
  return t.c[i];
}
int oracle_tableau_info(int method, int which) {
  Tab t = make_tab(method);
  if (t.S == 0) return -1;
  switch (which) {
    case 0: return t.S;
    case 1: return t.adaptive();
    case 2: return t.adaptive() ? t.fsal() : 0;
    case 3: return t.min_order();
  }
  return -1;
}
double oracle_pow(double x, double y, int mode) { return MathCtx{mode}.pow(x, y); }
double oracle_log10(double x, int mode) { return MathCtx{mode}.log10(x); }
void oracle_pow_v(const double* x, const double* y, double* o, int64_t n, int mode) {
  MathCtx M{mode};
  for (int64_t i = 0; i < n; i++) o[i] = M.pow(x[i], y[i]);
}
void oracle_log10_v(const double* x, double* o, int64_t n, int mode) {
  MathCtx M{mode};
  for (int64_t i = 0; i < n; i++) o[i] = M.log10(x[i]);
}
void oracle_rhs(int rhs, int64_t dof, const double* params, double t, const double* y, double* f) {
  Rhs F{rhs, dof, params};
  F(t, y, f);
}
double oracle_norm2(const double* x, int64_t n) { return norm2(x, n); }
void oracle_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#endif
}
int oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
}

/* ode_b200_detmath.h -- deterministic pow / log10 shared by host and device.
 *
 * Why this exists: the reference's step controllers call Base `^` and `log10`
 * (/root/reference src/runge_kutta.jl:436 `(1/err)^(1/(order+1))`,
 * src/ODE.jl:107-108 `10^(-(2+log10(..))/(p+1))`, src/ODE.jl:357
 * `(delta/err)^(1/3)`).  Julia's, glibc's and CUDA's libm differ at the ulp
 * level, and one ulp in `newdt` is enough to make a CUDA kernel and a CPU
 * checker diverge bit-wise.  This header is a single implementation built only
 * from IEEE-754 binary64 add/mul/fma (all correctly rounded on x86-64 and on
 * sm_100a) plus integer bit manipulation and two small tables, so the very same
 * bits come out on the CPU and on the GPU.  Its results are within
 * 0.5 ulp + ~2^-15 ulp of the exact value in the domain the solvers use
 * (|y*log(x)| < 64), i.e. they equal the correctly rounded result except in
 * rare near-tie cases (checked against mpmath in tests/test_detmath.py).
 *
 * The exponents the step controllers actually use -- the doubles nearest 1/3, 1/5 and 1/8 -- take a K-th root path
 * (dm_root3/5/8 below: the same argument reduction, a table of (2^s/invc)^(1/K), one short binomial series) that
 * costs half the arithmetic of the general log/exp path at the same accuracy; every other exponent (hinit's
 * 10^x, src/ODE.jl:108) takes the general path.
 *
 * Technique (standard, table-driven): log via an exact fma reduction
 * r = z*invc - 1 with an 8-bit invc and a 128-entry log(c) table kept as
 * hi + lo, a degree-10 Taylor tail (pairwise evaluation), result carried as a double-double; exp via
 * k = round(x*128/ln2), an exact hi reduction, a 128-entry 2^(j/128) table and a
 * final compensated `th + th*q + ...` so there is one effective rounding.
 *
 * Compile flags that matter: device code with -fmad=false, host code with
 * -ffp-contract=off, so that nothing except the explicit dm_fma calls fuses.
 */
#ifndef ODE_B200_DETMATH_H
#define ODE_B200_DETMATH_H

#ifdef __CUDACC_RTC__ /* NVRTC: no host headers; device code only */
#ifndef ODE_B200_H
typedef int int32_t;
typedef unsigned int uint32_t;
typedef long long int64_t;
typedef unsigned long long uint64_t;
#endif
#ifndef INFINITY
#define INFINITY __longlong_as_double(0x7ff0000000000000LL)
#endif
#ifndef NAN
#define NAN __longlong_as_double(0x7ff8000000000000LL)
#endif
#else
#include <stdint.h>
#include <string.h>
#include <math.h>
#endif

#include "ode_b200_detmath_tables.h"

#if defined(__CUDACC__)
#define DM_HD __host__ __device__ __forceinline__
#else
#define DM_HD static inline
#endif

/* ---- tables: one host copy, one device copy (same bits) ------------------ */
#define DM_X3(a, b, c) a, b, c,
#define DM_X2(a, b) a, b,
#ifndef __CUDACC_RTC__
static const double dm_logtab_h[128 * 3] = {DM_LOG_TABLE(DM_X3)};
static const double dm_exptab_h[128 * 2] = {DM_EXP_TABLE(DM_X2)};
static const double dm_root3tab_h[3 * 128 * 2] = {DM_ROOT3_TABLE(DM_X2)};
static const double dm_root5tab_h[5 * 128 * 2] = {DM_ROOT5_TABLE(DM_X2)};
static const double dm_root8tab_h[8 * 128 * 2] = {DM_ROOT8_TABLE(DM_X2)};
#endif
#if defined(__CUDACC__)
static __device__ const double dm_logtab_d[128 * 3] = {DM_LOG_TABLE(DM_X3)};
static __device__ const double dm_exptab_d[128 * 2] = {DM_EXP_TABLE(DM_X2)};
static __device__ const double dm_root3tab_d[3 * 128 * 2] = {DM_ROOT3_TABLE(DM_X2)};
static __device__ const double dm_root5tab_d[5 * 128 * 2] = {DM_ROOT5_TABLE(DM_X2)};
static __device__ const double dm_root8tab_d[8 * 128 * 2] = {DM_ROOT8_TABLE(DM_X2)};
#endif

#if defined(__CUDA_ARCH__)
#define DM_LOGTAB(i) dm_logtab_d[(i)]
#define DM_EXPTAB(i) dm_exptab_d[(i)]
#define DM_ROOTTAB(K, i) ((K) == 3 ? dm_root3tab_d[(i)] : ((K) == 5 ? dm_root5tab_d[(i)] : dm_root8tab_d[(i)]))
#else
#define DM_LOGTAB(i) dm_logtab_h[(i)]
#define DM_EXPTAB(i) dm_exptab_h[(i)]
#define DM_ROOTTAB(K, i) ((K) == 3 ? dm_root3tab_h[(i)] : ((K) == 5 ? dm_root5tab_h[(i)] : dm_root8tab_h[(i)]))
#endif

DM_HD uint64_t dm_bits(double x) {
#if defined(__CUDA_ARCH__)
  return (uint64_t)__double_as_longlong(x);
#else
  uint64_t u;
  memcpy(&u, &x, 8);
  return u;
#endif
}
DM_HD double dm_from_bits(uint64_t u) {
#if defined(__CUDA_ARCH__)
  return __longlong_as_double((long long)u);
#else
  double x;
  memcpy(&x, &u, 8);
  return x;
#endif
}
DM_HD double dm_fma(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
  return __fma_rn(a, b, c);
#else
  return __builtin_fma(a, b, c);
#endif
}

typedef struct {
  double hi, lo;
} dm_dd;

/* s + e = a + b exactly (Knuth two-sum, no magnitude precondition). */
DM_HD dm_dd dm_two_sum(double a, double b) {
  dm_dd r;
  double s = a + b;
  double bb = s - a;
  double e = (a - (s - bb)) + (b - bb);
  r.hi = s;
  r.lo = e;
  return r;
}

/* log(x) as a double-double for finite normal-or-subnormal x > 0. */
DM_HD dm_dd dm_log_dd(double x) {
  uint64_t ix = dm_bits(x);
  int64_t kadj = 0;
  if (ix < 0x0010000000000000ULL) { /* subnormal: scale by 2^52 (exact) */
    x = x * 4503599627370496.0;
    ix = dm_bits(x);
    kadj = -52;
  }
  const uint64_t OFF = 0x3FE6A00000000000ULL; /* 181/256 */
  uint64_t tmp = ix - OFF;
  int i = (int)((tmp >> 45) & 127);
  int64_t k = ((int64_t)tmp >> 52) + kadj;
  uint64_t iz = ix - (tmp & 0xFFF0000000000000ULL);
  double z = dm_from_bits(iz);
  double kd = (double)k;
  double invc = DM_LOGTAB(3 * i), logc = DM_LOGTAB(3 * i + 1), logct = DM_LOGTAB(3 * i + 2);

  double r = dm_fma(z, invc, -1.0);         /* exact */
  double t1 = dm_fma(kd, DM_LN2_HI, logc);  /* exact: both on the 2^-42 grid */
  /* fast two-sums are exact here: t1 is 0 or |t1| >= 1.9 |r| (checked over the whole table; |r| < 2^-7),
     and |t1 + r| >= r^2/2 */
  dm_dd a;
  a.hi = t1 + r;
  a.lo = (t1 - a.hi) + r;
  double r2 = r * r;
  double r2e = dm_fma(r, r, -r2);           /* r*r = r2 + r2e */
  double hr2 = -0.5 * r2;
  dm_dd b;
  b.hi = a.hi + hr2;                        /* ... - r^2/2 */
  b.lo = (a.hi - b.hi) + hr2;
  /* r^3/3 - r^4/4 + ... + r^9/9 - r^10/10 = r^3 * (c0 + c1 r + ... + c7 r^7), evaluated pairwise
     (Estrin) so the dependency chain is 3 fma deep instead of 7; the tail is < 2^-24 of the result */
  double p01 = dm_fma(-0.25, r, 0x1.5555555555555p-2);                  /* 1/3 - r/4 */
  double p23 = dm_fma(-0x1.5555555555555p-3, r, 0.2);                   /* 1/5 - r/6 */
  double p45 = dm_fma(-0.125, r, 0x1.2492492492492p-3);                 /* 1/7 - r/8 */
  double p67 = dm_fma(-0.1, r, 0x1.c71c71c71c71cp-4);                   /* 1/9 - r/10 */
  double r4 = r2 * r2;
  double p03 = dm_fma(p23, r2, p01);
  double p47 = dm_fma(p67, r2, p45);
  double p = dm_fma(p47, r4, p03);
  p = p * (r * r2);
  double lo = ((a.lo + b.lo) + (-0.5 * r2e)) + dm_fma(kd, DM_LN2_LO, logct);
  lo = lo + p;
  dm_dd res;
  res.hi = b.hi + lo;
  res.lo = (b.hi - res.hi) + lo; /* fast two-sum: |b.hi| >= |lo| or both tiny */
  return res;
}

/* exp(ehi + elo), |elo| << |ehi|.  The common case |ehi| < 700 is detected with one
 * integer compare on the high word and scales by a normal power of two. */
DM_HD double dm_exp_dd(double ehi, double elo) {
  const uint32_t hw = (uint32_t)(dm_bits(ehi) >> 32) & 0x7fffffffu;
  const int fast = hw < 0x4085E000u; /* |ehi| < 700 */
  if (!fast) {
    if (ehi != ehi) return ehi;
    if (ehi > 709.782712893384) return (double)INFINITY;
    if (ehi < -745.2) return 0.0;
  }
  double zz = ehi * DM_INVLN2N;
  const double SHIFT = 6755399441055744.0; /* 1.5 * 2^52 */
  double kd = (zz + SHIFT) - SHIFT;        /* round to nearest integer */
  int64_t ki = (int64_t)kd;
  double rh = dm_fma(kd, -DM_LN2N_HI, ehi); /* exact (see header comment) */
  double rl = dm_fma(kd, -DM_LN2N_LO, elo);
  dm_dd qq = dm_two_sum(rh, rl);
  double q = qq.hi, ql = qq.lo;
  int j = (int)(ki & 127);
  int64_t e2 = (ki - j) / 128; /* exact: ki - j is a multiple of 128 */
  double th = DM_EXPTAB(2 * j), tl = DM_EXPTAB(2 * j + 1);
  /* exp(q) - 1 - q = q^2 (1/2 + q/6 + ... + q^5/5040) */
  /* pairwise (Estrin): 3 fma deep instead of 6 */
  double q2 = q * q;
  double e01 = dm_fma(0x1.5555555555555p-3, q, 0.5);                     /* 1/2 + q/6 */
  double e23 = dm_fma(0x1.1111111111111p-7, q, 0x1.5555555555555p-5);    /* 1/24 + q/120 */
  double e45 = dm_fma(0x1.a01a01a01a01ap-13, q, 0x1.6c16c16c16c17p-10);  /* 1/720 + q/5040 */
  double q4 = q2 * q2;
  double p = dm_fma(e45, q4, dm_fma(e23, q2, e01));
  double rest = dm_fma(q2, p, dm_fma(q, ql, ql));
  double ph = th * q;
  double pl = dm_fma(th, q, -ph);
  double s = th + ph;
  double e = (th - s) + ph; /* fast two-sum, |th| >= |ph| */
  double tail = (e + pl) + dm_fma(th, rest, dm_fma(tl, q, tl));
  double v = s + tail; /* in [1, 2) roughly; scale by 2^e2 */
  if (fast || (e2 >= -1021 && e2 <= 1022)) {
    return v * dm_from_bits((uint64_t)(e2 + 1023) << 52);
  }
  /* out-of-range scale: two exact-power-of-two multiplies (deterministic) */
  if (e2 > 1022) {
    double f = dm_from_bits((uint64_t)(1022 + 1023) << 52);
    return (v * f) * dm_from_bits((uint64_t)(e2 - 1022 + 1023) << 52);
  } else {
    double f = dm_from_bits((uint64_t)(-1021 + 1023) << 52);
    int64_t rem = e2 + 1021; /* negative */
    if (rem < -1021) rem = -1021;
    return (v * f) * dm_from_bits((uint64_t)(rem + 1023) << 52);
  }
}

/* x^y for the three exponents the step controllers use, y = the double nearest 1/K with K = 3, 5, 8
 * ((1/err)^(1/(order+1)), src/runge_kutta.jl:436; (delta/err)^(1/3), src/ODE.jl:357), for positive finite x
 * (normal or subnormal).  Half the arithmetic of the general path and the same accuracy (0.5 ulp + ~2^-15 ulp):
 *     x = 2^k z,  z*invc_i = 1 + r exactly (the log table's reduction),  k = K q + s,  0 <= s < K
 *     x^(1/K) = 2^q * R[s][i] * (1 + r)^(1/K)            R[s][i] = (2^s / invc_i)^(1/K) tabulated as hi + lo
 *     (1 + r)^(1/K) = 1 + r/K + r^2 Q(r)                 |r| < 2^-7: binomial series to r^9, r/K with 1/K = AH + AL
 *     x^y = x^(1/K) * exp((y - 1/K) ln x)                y - 1/K = DELTA ~ 1e-17: 1 + DELTA ln x, ln x to 2^-40
 * with the sum R*(1 + ...) accumulated so that there is one effective rounding. */
#define DM_ROOT_BODY(K, AH, AL, DELTA, C0, C1, C2, C3, C4, C5, C6, C7)                                              \
  uint64_t ix = dm_bits(x);                                                                                         \
  int64_t kadj = 0;                                                                                                 \
  if (ix < 0x0010000000000000ULL) { /* subnormal: scale by 2^52 (exact) */                                          \
    x = x * 4503599627370496.0;                                                                                     \
    ix = dm_bits(x);                                                                                                \
    kadj = -52;                                                                                                     \
  }                                                                                                                 \
  const uint64_t tmp = ix - 0x3FE6A00000000000ULL;                                                                  \
  const int i = (int)((tmp >> 45) & 127);                                                                           \
  const int64_t k = ((int64_t)tmp >> 52) + kadj;                                                                    \
  const double z = dm_from_bits(ix - (tmp & 0xFFF0000000000000ULL));                                                \
  int64_t sft = k % (K);                                                                                            \
  if (sft < 0) sft += (K);                                                                                          \
  const int64_t q = (k - sft) / (K);                                                                                \
  const double invc = DM_LOGTAB(3 * i), logc = DM_LOGTAB(3 * i + 1);                                                \
  const double rh = DM_ROOTTAB(K, 2 * ((int)sft * 128 + i)), rl = DM_ROOTTAB(K, 2 * ((int)sft * 128 + i) + 1);      \
  const double r = dm_fma(z, invc, -1.0); /* exact */                                                               \
  const double t1h = r * (AH);                                                                                      \
  double plo = dm_fma(r, (AL), dm_fma(r, (AH), -t1h)); /* r/K = t1h + plo */                                        \
  const double r2 = r * r;                                                                                          \
  const double p01 = dm_fma((C1), r, (C0)), p23 = dm_fma((C3), r, (C2));                                            \
  const double p45 = dm_fma((C5), r, (C4)), p67 = dm_fma((C7), r, (C6));                                            \
  const double r4 = r2 * r2;                                                                                        \
  const double qq = dm_fma(dm_fma(p67, r2, p45), r4, dm_fma(p23, r2, p01));                                         \
  plo = dm_fma(r2, qq, plo);                                                                                        \
  if ((DELTA) != 0.0) { /* (1 + t1h + plo)(1 + c) = 1 + t1h + (plo + c + c*t1h), c = DELTA ln x */                  \
    const double c = (DELTA) * (dm_fma((double)k, DM_LN2_HI, logc) + r);                                            \
    plo = dm_fma(c, t1h, plo) + c;                                                                                  \
  }                                                                                                                 \
  const double w = rh * t1h;                                                                                        \
  const double v = dm_fma(rh, plo, dm_fma(rl, t1h, rl)) + dm_fma(rh, t1h, -w);                                      \
  const double sm = rh + w;                                                                                         \
  const double tail = ((rh - sm) + w) + v; /* fast two-sum: |rh| > |w| */                                           \
  return (sm + tail) * dm_from_bits((uint64_t)(q + 1023) << 52);

DM_HD double dm_root3(double x) {
  DM_ROOT_BODY(3, DM_ROOT3_AH, DM_ROOT3_AL, DM_ROOT3_DELTA, DM_ROOT3_C0, DM_ROOT3_C1, DM_ROOT3_C2, DM_ROOT3_C3, DM_ROOT3_C4,
               DM_ROOT3_C5, DM_ROOT3_C6, DM_ROOT3_C7)
}
DM_HD double dm_root5(double x) {
  DM_ROOT_BODY(5, DM_ROOT5_AH, DM_ROOT5_AL, DM_ROOT5_DELTA, DM_ROOT5_C0, DM_ROOT5_C1, DM_ROOT5_C2, DM_ROOT5_C3, DM_ROOT5_C4,
               DM_ROOT5_C5, DM_ROOT5_C6, DM_ROOT5_C7)
}
DM_HD double dm_root8(double x) {
  DM_ROOT_BODY(8, DM_ROOT8_AH, DM_ROOT8_AL, DM_ROOT8_DELTA, DM_ROOT8_C0, DM_ROOT8_C1, DM_ROOT8_C2, DM_ROOT8_C3, DM_ROOT8_C4,
               DM_ROOT8_C5, DM_ROOT8_C6, DM_ROOT8_C7)
}

/* x^y for x >= 0.  Negative x returns NaN (Julia raises DomainError there;
 * the solvers only ever pass norms and ratios of norms).  Positive normal
 * finite x with finite y -- every controller call -- is recognised with integer
 * compares only and goes straight to the arithmetic. */
DM_HD double dm_pow(double x, double y) {
  const uint64_t ix = dm_bits(x);
  const uint64_t iy = dm_bits(y) & 0x7fffffffffffffffULL;
  const int xnorm = (ix - 0x0010000000000000ULL) < 0x7fe0000000000000ULL; /* +normal, finite */
  const int yfin = iy < 0x7ff0000000000000ULL;
  if (!(xnorm && yfin)) {
    if (x != x || y != y) return x + y; /* NaN */
    if (y == 0.0) return 1.0;
    if (x < 0.0) return (double)NAN;
    if (x == 0.0) return y > 0.0 ? 0.0 : (double)INFINITY;
    if (x == (double)INFINITY) return y > 0.0 ? (double)INFINITY : 0.0;
    if (y == (double)INFINITY) return x == 1.0 ? 1.0 : (x > 1.0 ? (double)INFINITY : 0.0);
    if (y == -(double)INFINITY) return x == 1.0 ? 1.0 : (x > 1.0 ? 0.0 : (double)INFINITY);
    /* positive subnormal x, finite y: falls through */
  }
#ifndef DM_NO_ROOTS
  /* the controllers' exponents (compile-time constants at every hot call site: the tests fold away) */
  if (y == 0.5) return sqrt(x); /* IEEE square root on both sides: the correctly rounded x^(1/2) */
  if (y == DM_ROOT5_AH) return dm_root5(x);
  if (y == DM_ROOT3_AH) return dm_root3(x);
  if (y == DM_ROOT8_AH) return dm_root8(x);
#endif
  dm_dd l = dm_log_dd(x);
  double ehi = y * l.hi;
  double elo = dm_fma(y, l.lo, dm_fma(y, l.hi, -ehi));
  return dm_exp_dd(ehi, elo);
}

/* log10(x) for x >= 0. */
DM_HD double dm_log10(double x) {
  if (x != x) return x;
  if (x < 0.0) return (double)NAN;
  if (x == 0.0) return -(double)INFINITY;
  if (x == (double)INFINITY) return x;
  dm_dd l = dm_log_dd(x);
  double ph = l.hi * DM_INVLN10_HI;
  double pe = dm_fma(l.hi, DM_INVLN10_HI, -ph);
  double t = dm_fma(l.hi, DM_INVLN10_LO, l.lo * DM_INVLN10_HI);
  return ph + (pe + t);
}

#endif /* ODE_B200_DETMATH_H */

// ensemble_adapt.cuh -- adaptive embedded Runge-Kutta, one trajectory per thread.
//
// Replaces n_traj independent calls of the reference driver
//   oderk_adapt         /root/reference src/runge_kutta.jl:232-369
//   rk_embedded_step!   :371-399      calc_next_k!  :444-458
//   stepsize_hw92!      :401-442      hinit         src/ODE.jl:85-112
//   hermite_interp!     :479-492
// with identical arithmetic (binary64, no FMA, reference association order).
//
// Mapping to the machine
//  * state y, k1, the S-1 scaled stages dt*k_s and both b-row accumulators live
//    in registers; tableau coefficients are compile-time constants (zero entries
//    drop out of the unrolled loops); tolerances come straight from the
//    kernel-parameter constant bank.
//  * dt*k_s is formed once per stage: (dt*k)*a is left-associated in the
//    reference (:454), so hoisting dt*k is bit-identical and removes
//    nnz(a) - (S-1) multiplies per attempt.
//  * the grid is persistent: every lane pulls trajectory indices from a global
//    counter, so a warp does not idle until its slowest trajectory ends; the
//    loop body is one step ATTEMPT, common to all lanes, and accept/reject is
//    a handful of predicated register moves (FSAL) rather than a branch around
//    the stage loop.
//  * zero tableau coefficients are dropped at compile time.  That is bit-identical to the reference
//    (which multiplies them, :385-392,454) as long as every stage is finite; with a non-finite stage
//    the reference forms Inf*0 = NaN.  The fast kernel therefore only DETECTS a non-finite trial
//    state / scaled error (integer tests), marks the trajectory ODE_B200_ST_REDO and drops it; a
//    second launch of the LITERAL instantiation (every coefficient multiplied, zeros included)
//    re-integrates just those trajectories from t0, so blow-ups follow the reference's path
//    (NaN guard -> dt/5 -> ... -> minstep stop) exactly, at no cost to the hot loop.
//  * MODE 0 keeps only final state + counters; MODE 1 adds the reference's
//    tout/yout contract (points = :all / :specified with Hermite output) and an
//    optional attempt trace, used by the odeXX front-ends and the parity tests.
#pragma once

#include "common.cuh"
#include "rhs.cuh"
#include "tableaus.cuh"

namespace odeb200 {

template <int D>
__device__ __forceinline__ void hermite_point(double* out, double tq, double t, double dt, const double (&y0)[D],
                                              const double (&y1)[D], const double (&f0)[D], const double (&f1)[D]) {
  double theta = (tq - t) / dt;  // :486
#pragma unroll
  for (int i = 0; i < D; i++) {  // :488-489
    out[i] = ((1.0 - theta) * y0[i] + theta * y1[i]) +
             (theta * (theta - 1.0)) *
                 (((1.0 - 2.0 * theta) * (y1[i] - y0[i]) + ((theta - 1.0) * dt) * f0[i]) + (theta * dt) * f1[i]);
  }
}

// hinit, src/ODE.jl:85-112 with p = min(order): returns the signed first step, leaves f0 in k1 (-> ks[1]).
template <class Tab, class Rhs>
__device__ __forceinline__ double hinit_rk(const EnsembleArgs& A, double tstart, double tend, double tdir,
                                           const double (&y)[Rhs::DOF], const double (&p)[RhsNP<Rhs>::value],
                                           double (&k1)[Rhs::DOF]) {
  constexpr int D = Rhs::DOF;
  double n0 = fabs(y[0]);
#pragma unroll
  for (int d = 1; d < D; d++) n0 = normInf_step(n0, fabs(y[d]));
  double tau = jl_max(A.reltol * n0, A.abstol);  // :89
  double d0 = n0 / tau;                          // :90
  Rhs::eval(tstart, y, k1, p);                   // :91  f0
  double n1 = fabs(k1[0]);
#pragma unroll
  for (int d = 1; d < D; d++) n1 = normInf_step(n1, fabs(k1[d]));
  double d1 = n1 / tau;  // :92
  double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * (d0 / d1);  // :93-97
  double x1[D], f1[D];
  double th = tdir * h0;
#pragma unroll
  for (int d = 0; d < D; d++) x1[d] = y[d] + th * k1[d];  // :100
  Rhs::eval(tstart + tdir * h0, x1, f1, p);               // :101
  double n2 = fabs(f1[0] - k1[0]);
#pragma unroll
  for (int d = 1; d < D; d++) n2 = normInf_step(n2, fabs(f1[d] - k1[d]));
  double d2 = n2 / (tau * h0);  // :103
  double dm = jl_max(d1, d2);
  double h1;
  if (dm <= 1e-15) {
    h1 = jl_max(1e-6, 1e-3 * h0);  // :105
  } else {
    double pw = -(2.0 + dm_log10(dm)) / (double)(Tab::ORDER + 1);  // :107
    h1 = dm_pow(10.0, pw);                                         // :108
  }
  return tdir * jl_min(jl_min(100.0 * h0, h1), tdir * (tend - tstart));  // :111
}

// Initial step and f0 of every trajectory, one lane per trajectory with full warps.  Inside the
// persistent kernel a lane that fetches a new trajectory would run these ~300 instructions alone
// while its 31 neighbours wait (2.6 % of all issue slots on Lorenz/dopri5, profiles/ensemble_lorenz_r1.md);
// here the same arithmetic costs 1/32 of that.  init[idx] = {dt, f0[0..D)}.
template <class Tab, class Rhs>
__global__ void __launch_bounds__(ENS_BLOCK) ens_adapt_init_kernel(const EnsembleArgs A) {
  constexpr int D = Rhs::DOF, NPA = RhsNP<Rhs>::value;
  const double tstart = A.tspan[0];
  const double tend = A.tspan[A.n_tspan - 1];
  const double tdir = jl_sign(tend - tstart);
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < A.n_traj;
       idx += (long long)gridDim.x * blockDim.x) {
    double y[D], k1[D], p[NPA];
#pragma unroll
    for (int d = 0; d < D; d++) y[d] = A.y0[idx * D + d];
#pragma unroll
    for (int q = 0; q < NPA; q++) p[q] = (q < Rhs::NP) ? A.params[idx * A.params_stride + q] : 0.0;
    double* out = A.init + idx * (D + 1);
    out[0] = hinit_rk<Tab, Rhs>(A, tstart, tend, tdir, y, p, k1);
#pragma unroll
    for (int d = 0; d < D; d++) out[1 + d] = k1[d];
  }
}

// The same pre-pass reading y0 straight from the caller's page-locked host buffer (device-mapped), in trajectory
// order, WHILE the persistent kernel runs: the upload of y0 (25 MB = 0.45 ms of an 11.3 ms call for 2^20 Lorenz
// trajectories) disappears behind the integration.  A FEW fat blocks (STREAM_BLOCKS x 256 threads: enough loads in
// flight to keep PCIe busy), launched before the persistent grid so that they are resident first; each walks the
// chunks c0 + blockIdx, c0 + blockIdx + gridDim, ... < c1 and raises ready[chunk] when the chunk's rows {dt0, f0, y0}
// are in HBM.  A lane of the persistent kernel that fetches a trajectory whose chunk is not ready yet waits for the
// flag; the pre-pass never waits for anything, so the pair cannot deadlock (and a lane that waited too long computes
// hinit itself from the caller's buffer).  Only for one parameter vector shared by the whole ensemble.
// Measured alternatives (profiles/ensemble_kernels_r2.md): one 64-thread block on EVERY SM costs each SM a resident
// CTA of the persistent grid for 0.5 ms (register files are per scheduler: a 64-register warp does not fit next to
// five 96-register ones), 11.30 -> 11.16 ms only; pre-processing a DMA-copied remainder with a second launch cannot
// become resident next to the full persistent grid at all (the lanes time out: 290 ms).
template <class Tab, class Rhs>
__global__ void __launch_bounds__(STREAM_CHUNK, 2) ens_adapt_stream_init_kernel(const EnsembleArgs A, int* ready, long long cfirst, long long cend, int nlong) {
  constexpr int D = Rhs::DOF, NPA = RhsNP<Rhs>::value, W = 2 * D + 1;
  constexpr int PAIRS = STREAM_CHUNK * D / 2;                          // 16-byte pieces of one chunk of y0
  constexpr int PER = (PAIRS + STREAM_CHUNK - 1) / STREAM_CHUNK;       // pieces per thread
  static_assert((STREAM_CHUNK * D) % 2 == 0, "a chunk of y0 is a whole number of 16-byte pieces");
  __shared__ double2 sy[PAIRS];
  const double tstart = A.tspan[0];
  const double tend = A.tspan[A.n_tspan - 1];
  const double tdir = jl_sign(tend - tstart);
  double p[NPA];
#pragma unroll
  for (int q = 0; q < NPA; q++) p[q] = (q < Rhs::NP) ? A.uparams[q] : 0.0;
  // Blocks 0 .. cfirst-1 take one chunk each first -- the rows every resident lane of the persistent grid starts with,
  // all in flight at once -- and all but the first `nlong` blocks leave after that (making room for the persistent
  // grid); the long-lived ones share the remaining chunks round-robin.
  const long long total = A.n_traj * D;  // doubles in y0
  auto next_chunk = [&](long long c) -> long long {
    if (c < cfirst) return (blockIdx.x < nlong) ? cfirst + blockIdx.x : cend;
    return c + nlong;
  };
  // The chunk's 64*D doubles are fetched over PCIe as coalesced 16-byte loads (a warp asks for 512 contiguous bytes),
  // one chunk ahead of the arithmetic, and handed to their trajectories through shared memory.
  double2 nx[PER];
  auto fetch = [&](long long c) {
    const long long base = c * (STREAM_CHUNK * D);
#pragma unroll
    for (int k = 0; k < PER; k++) {
      const int i = threadIdx.x + k * STREAM_CHUNK;
      const long long g = base + 2 * i;
      nx[k] = make_double2(0.0, 0.0);
      if (i < PAIRS && g + 1 < total) nx[k] = *reinterpret_cast<const double2*>(A.y0 + g);
      else if (i < PAIRS && g < total) nx[k].x = A.y0[g];
    }
  };
  long long c = (blockIdx.x < cfirst) ? (long long)blockIdx.x : ((blockIdx.x < nlong) ? cfirst + blockIdx.x : cend);
  if (c < cend) fetch(c);
  for (; c < cend; c = next_chunk(c)) {
#pragma unroll
    for (int k = 0; k < PER; k++) {
      const int i = threadIdx.x + k * STREAM_CHUNK;
      if (i < PAIRS) sy[i] = nx[k];
    }
    __syncthreads();
    if (next_chunk(c) < cend) fetch(next_chunk(c));
    const long long idx = c * STREAM_CHUNK + threadIdx.x;
    if (idx < A.n_traj) {
      double y[D], k1[D];
      const double* mine = reinterpret_cast<const double*>(sy) + threadIdx.x * D;
#pragma unroll
      for (int d = 0; d < D; d++) y[d] = mine[d];
      double* out = A.init + idx * W;
      out[0] = hinit_rk<Tab, Rhs>(A, tstart, tend, tdir, y, p, k1);
#pragma unroll
      for (int d = 0; d < D; d++) {
        out[1 + d] = k1[d];
        out[1 + D + d] = y[d];
      }
    }
    __threadfence();
    __syncthreads();  // (also: everybody has read sy before the next chunk overwrites it)
    if (threadIdx.x == 0) *(volatile int*)&ready[c] = 1;
  }
}

// Resident CTAs per SM asked of ptxas.  Measured on Lorenz/dopri5 (profiles/ensemble_lorenz_r1.md):
// 5 CTAs/SM (<= 102 registers, no spill) beats both the unconstrained allocation (119 registers,
// 4 CTAs/SM) and tighter bounds that spill.  Larger state x stage products are left unconstrained.
template <class Tab, class Rhs>
constexpr int adapt_min_blocks() {
  // 13 stages x 4 components: unconstrained ptxas takes 152 registers (3 CTAs/SM); 128 registers with a 96-byte spill
  // (4 CTAs/SM) measured 2 % faster on Kepler/feh78 (17.66 -> 17.29 ms, profiles/ensemble_kernels_r2.md)
  if (ODEB200_WIDE_MINB == 0 && Rhs::DOF >= 4 && Tab::S >= 13 && ENS_BLOCK == 128) return 4;
#ifdef ODEB200_NARROW_MINB
  return (Rhs::DOF <= 3 && Rhs::DOF * Tab::S <= 21) ? ODEB200_NARROW_MINB : ODEB200_WIDE_MINB;
#else
  return (Rhs::DOF <= 3 && Rhs::DOF * Tab::S <= 21) ? (5 * 128) / ENS_BLOCK : ODEB200_WIDE_MINB;  // 640 resident threads per SM; 0 = leave it to ptxas
#endif
}

// Which (tableau, functor) pairs take the deferred-check controller (ODEB200_CTRL_NB = 2), from measurement
// (profiles/ensemble_kernels_r2.md, 2^18 trajectories): Lorenz/dopri5 3.31 -> 3.03 ms, Kepler/feh78 18.22 -> 17.56 ms,
// but Lorenz/feh78 2.08 -> 2.14 ms and Van der Pol/dopri5 1.73 -> 1.79 ms: it pays where there are at least three
// divisions to straighten and the stage loop leaves registers for the kept numerators.
template <class Tab, class Rhs>
__host__ __device__ constexpr bool ctrl_nb_policy() {
  return Rhs::DOF >= 3 && (Tab::S <= 7 || Rhs::DOF >= 4);
}

template <class Tab, class Rhs, int MODE, bool LITERAL, bool PU = false>
__global__ void __launch_bounds__(ENS_BLOCK, LITERAL ? 1 : adapt_min_blocks<Tab, Rhs>() + (PU ? ODEB200_PU_EXTRA : 0))
    ens_adapt_kernel(const __grid_constant__ EnsembleArgs A) {
  if (LITERAL && *A.redo_count == 0) return;  // nothing was marked by the fast pass
  constexpr int D = Rhs::DOF, S = Tab::S, NPA = RhsNP<Rhs>::value;
  constexpr bool FSAL = Tab::fsal();
  constexpr double EXPO = 1.0 / (double)(Tab::ORDER + 1);  // T(1/(order+1)) :436
  constexpr int NK = S > 1 ? S - 1 : 1;
  constexpr bool BANK = ODEB200_TAB_BANK != 0;

  const double tstart = A.tspan[0];
  const double tend = A.tspan[A.n_tspan - 1];
  const double tdir = jl_sign(tend - tstart);  // src/ODE.jl:87 (zero span rejected on the host)
  const bool fwd = __double2hiint(tdir) > 0;    // tdir = +-1: x*tdir == (fwd ? x : -x) exactly (sign test off the FP64 pipe)
  const int nfixed = A.n_tspan;

  // Tail hand-off.  When the work counter runs out every lane is busy with its last trajectory and the lanes of a
  // warp finish at times spread over one trajectory duration: a warp would keep issuing for its slowest lane with
  // ever fewer lanes active (4-5 % of all warp-iterations on 2^20 Lorenz trajectories, tools/tail_sim.py).  Instead the
  // first lane of a warp that finds no work raises the warp's flag; its neighbours, at the end of their current
  // attempt, write their state to A.dump and leave; a second launch of this kernel (A.resume) continues those
  // trajectories, packed so that lanes of a warp have about the same distance to go.  The arithmetic of a trajectory
  // does not change: its state crosses the launches as the very same doubles.
  constexpr bool DRAIN = (ODEB200_DRAIN != 0) && MODE == 0 && !LITERAL;
  __shared__ int warp_out[DRAIN ? ENS_BLOCK / 32 : 1];
  if (DRAIN) {
    if (threadIdx.x < ENS_BLOCK / 32) warp_out[threadIdx.x] = 0;
    __syncthreads();
  }

  long long idx = -1;
  double y[D], k1[D], pl[NPA];
  // PU: one parameter vector for the whole ensemble, read from the kernel-parameter constant bank
  const double (&p)[NPA] = PU ? reinterpret_cast<const double (&)[NPA]>(A.uparams) : const_cast<const double (&)[NPA]>(pl);
  double t = 0.0, dt = 0.0;
  int timeout = 0, nacc = 0, nrej = 0;
  bool islast = false;
  int iter = 2, iter_fixed = 2, npts = 0;  // 1-based like the reference (:286,:304)
  long long ntrace = 0;

  while (true) {
    if (DRAIN && idx < 0 && A.resume != nullptr) {  // second launch: continue a handed-over trajectory
      const unsigned long long slot = atomicAdd(A.work_counter, 1ULL);
      if (slot >= (unsigned long long)*A.dump_count) break;
      const double* rec = A.resume + (long long)A.resume_order[slot] * dump_width(D);
      t = rec[1];
      dt = rec[2];
      idx = __double_as_longlong(rec[3]);
      const unsigned long long cnt = (unsigned long long)__double_as_longlong(rec[4]);
      const unsigned long long fl = (unsigned long long)__double_as_longlong(rec[5]);
      nacc = (int)(cnt & 0xffffffffULL);
      nrej = (int)(cnt >> 32);
      timeout = (int)(fl & 0xffULL);
      islast = ((fl >> 8) & 1ULL) != 0;
#pragma unroll
      for (int d = 0; d < D; d++) {
        y[d] = rec[6 + d];
        k1[d] = rec[6 + D + d];
      }
      if (!PU) {
#pragma unroll
        for (int q = 0; q < NPA; q++) pl[q] = (q < Rhs::NP) ? A.params[idx * A.params_stride + q] : 0.0;
      }
    }
    if (idx < 0) {
      idx = (long long)atomicAdd(A.work_counter, 1ULL);
      if (idx >= A.n_traj) {
        if (DRAIN && A.dump != nullptr) *(volatile int*)&warp_out[threadIdx.x >> 5] = 1;
        break;
      }
      if (!LITERAL && A.order != nullptr) idx = A.order[idx];  // longest expected trajectory first (common.cuh)
      if (LITERAL && A.status[idx] != ODE_B200_ST_REDO) {  // the literal pass redoes marked trajectories only
        idx = -1;
        continue;
      }
      // ---------------------------------------------------------- init :256-304
      bool have_row = false;
      if (MODE == 0 && !LITERAL && PU && A.ready != nullptr) {  // streaming pre-pass: {dt0, f0, y0} when the chunk is ready
        const volatile int* flag = A.ready + idx / STREAM_CHUNK;
        for (int spin = 0; spin < (1 << 16) && !have_row; spin++) {
          have_row = *flag != 0;
          if (!have_row) __nanosleep(200);
        }
        if (have_row) {
          __threadfence();
          const double* in = A.init + idx * (2 * D + 1);
          dt = __ldcg(in);
#pragma unroll
          for (int d = 0; d < D; d++) {
            k1[d] = __ldcg(in + 1 + d);
            y[d] = __ldcg(in + 1 + D + d);
          }
        }
      }
      if (!have_row) {
#pragma unroll
        for (int d = 0; d < D; d++) y[d] = A.y0[idx * D + d];
      }
      if (!PU) {
#pragma unroll
        for (int q = 0; q < NPA; q++) pl[q] = (q < Rhs::NP) ? A.params[idx * A.params_stride + q] : 0.0;
      }
      t = tstart;
      nacc = nrej = 0;
      timeout = 0;  // :303
      iter = 2;
      iter_fixed = 2;
      ntrace = 0;
      if (have_row) {
      } else if (MODE == 0 && !LITERAL && A.init != nullptr && A.ready == nullptr) {  // computed by ens_adapt_init_kernel with all lanes busy
        const double* in = A.init + idx * (D + 1);
        dt = in[0];
#pragma unroll
        for (int d = 0; d < D; d++) k1[d] = in[1 + d];
      } else {
        dt = hinit_rk<Tab, Rhs>(A, tstart, tend, tdir, y, p, k1);
      }
      if (A.initstep != 0.0) dt = A.initstep;            // :293-295 (sign checked on the host)
      islast = fabs(t + dt - tend) <= jl_eps(tend);      // :302
      if (MODE == 1) {
        if (A.points == ODE_B200_POINTS_ALL) {  // ys[1] = y0, tspan = [tstart] :280,:285
          if (A.pts_cap > 0) {
            A.pts_t[idx * A.pts_cap] = tstart;
#pragma unroll
            for (int d = 0; d < D; d++) A.pts_y[(idx * A.pts_cap) * D + d] = y[d];
          }
          npts = 1;
        } else if (A.points == ODE_B200_POINTS_SPECIFIED) {
          for (int i = 0; i < nfixed && i < A.pts_cap; i++) A.pts_t[idx * A.pts_cap + i] = A.tspan[i];
          if (A.pts_cap > 0) {
#pragma unroll
            for (int d = 0; d < D; d++) A.pts_y[(idx * A.pts_cap) * D + d] = y[d];
          }
          npts = nfixed;
        }
      }
    }

    // ------------------------------------------------ rk_embedded_step! :371-399
    double ytrial[D], yerr[D], dtk[NK][D], klast[D];
    bool rhs_bad = false;  // a deferred sqrt / division check of the right-hand side failed (rhs.cuh eval_nb)
    constexpr double b10 = Tab::b(0, 0), b20 = Tab::b(1, 0);
    const double b10v = BANK ? A.tab[tab_b_index(S, 0, 0)] : b10, b20v = BANK ? A.tab[tab_b_index(S, 1, 0)] : b20;
#pragma unroll
    for (int d = 0; d < D; d++) {  // :384-387
      ytrial[d] = LITERAL ? (0.0 + b10v * k1[d]) : ((b10 != 0.0) ? b10v * k1[d] : 0.0);
      yerr[d] = LITERAL ? (0.0 + b20v * k1[d]) : ((b20 != 0.0) ? b20v * k1[d] : 0.0);
      dtk[0][d] = dt * k1[d];
      klast[d] = k1[d];
    }
    static_for<S - 1>([&](auto I) {  // :388-394
      constexpr int s = decltype(I)::value + 1;
      double ytmp[D], ks[D];
#pragma unroll
      for (int d = 0; d < D; d++) ytmp[d] = y[d];  // calc_next_k! :452
      static_for<s>([&](auto J) {                  // :453-455
        constexpr int ss = decltype(J)::value;
        constexpr double a = Tab::a(s, ss);
        if constexpr (a != 0.0 || LITERAL) {
          const double av = BANK ? A.tab[tab_a_index(s, ss)] : a;  // same value: the host fills tab from this tableau
#pragma unroll
          for (int d = 0; d < D; d++) ytmp[d] = ytmp[d] + dtk[ss][d] * av;
        }
      });
      constexpr double c = Tab::c(s);
      const double cv = BANK ? A.tab[tab_c_index(S, s)] : c;
      if constexpr (LITERAL) Rhs::eval(t + cv * dt, ytmp, ks, p);  // :456
      else rhs_eval_fast<Rhs>(t + cv * dt, ytmp, ks, p, rhs_bad);  // same values; failed fast-path checks -> rhs_bad
      constexpr double b1 = Tab::b(0, s), b2 = Tab::b(1, s);
      const double b1v = BANK ? A.tab[tab_b_index(S, 0, s)] : b1, b2v = BANK ? A.tab[tab_b_index(S, 1, s)] : b2;
#pragma unroll
      for (int d = 0; d < D; d++) {  // :390-393
        if constexpr (b1 != 0.0 || LITERAL) ytrial[d] = ytrial[d] + b1v * ks[d];
        if constexpr (b2 != 0.0 || LITERAL) yerr[d] = yerr[d] + b2v * ks[d];
        if constexpr (s < S - 1) dtk[s][d] = dt * ks[d];
        if constexpr (s == S - 1) klast[d] = ks[d];
      }
    });
#pragma unroll
    for (int d = 0; d < D; d++) {  // :395-398
      yerr[d] = dt * (ytrial[d] - yerr[d]);
      ytrial[d] = y[d] + dt * ytrial[d];
    }

    // ---------------------------------------------------- stepsize_hw92! :401-442
    bool nanout = false, weird = false;
    unsigned int expo_or = 0u;  // fast pass: OR of the high words of every trial component and scaled error
    // ODEB200_CTRL_NB: the D divisions of :433, the square root of :435 and 1/err of :436 as straight-line code with
    // deferred acceptance checks (common.cuh div_nb / sqrt_nb): one rarely taken branch that redoes them with the
    // compiler's own operations instead of a slow-path branch (and a basic-block boundary) per operation
    constexpr bool CNB = !LITERAL && (ODEB200_CTRL_NB == 1 || (ODEB200_CTRL_NB == 2 && ctrl_nb_policy<Tab, Rhs>()));
    bool cbad = false;
    double ynum[CNB ? D : 1];
#pragma unroll
    for (int d = 0; d < D; d++) {  // :430-434
      if (LITERAL) {
        const double ay = fabs(y[d]), at = fabs(ytrial[d]);
        nanout = nanout || (ytrial[d] != ytrial[d]);  // isoutofdomain, src/ODE.jl:116
        yerr[d] = yerr[d] / (A.abstol + jl_max(ay, at) * A.reltol);  // :433
      } else if (CNB) {
        // max(|y|, |ytrial|) as an integer compare of the magnitudes' bit patterns (order-preserving for non-NaN values)
        const double ay = fabs(y[d]), at = fabs(ytrial[d]);
        const bool ygt = (unsigned long long)__double_as_longlong(ay) > (unsigned long long)__double_as_longlong(at);
        ynum[d] = yerr[d];
        yerr[d] = div_nb(yerr[d], recip_prepare(A.abstol + (ygt ? ay : at) * A.reltol), cbad);
        const unsigned int ht = (unsigned int)__double2hiint(ytrial[d]) & 0x7ff00000u;
        expo_or = expo_or > ht ? expo_or : ht;
      } else {
        // max(|y|, |ytrial|) as an integer compare of the magnitudes' bit patterns (order-preserving for
        // non-NaN values; a NaN trial state sends the trajectory to the literal pass whatever is picked here)
        const double ay = fabs(y[d]), at = fabs(ytrial[d]);
        const bool ygt = (unsigned long long)__double_as_longlong(ay) > (unsigned long long)__double_as_longlong(at);
        yerr[d] = yerr[d] / (A.abstol + (ygt ? ay : at) * A.reltol);  // :433
        // non-finite <=> exponent field all ones; testing max(exponent fields) keeps these 2*D tests off the
        // FP64 pipe (written on the high words so that nvcc does not turn them back into DSETP)
        const unsigned int ht = (unsigned int)__double2hiint(ytrial[d]) & 0x7ff00000u;
        const unsigned int he = (unsigned int)__double2hiint(yerr[d]) & 0x7ff00000u;
        expo_or = expo_or > ht ? expo_or : ht;
        expo_or = expo_or > he ? expo_or : he;
      }
    }
    double inv_err = 0.0, err_nb = 0.0;
    if (CNB) {
      err_nb = norm2_nb<D>(yerr, cbad);
      inv_err = div_nb(1.0, recip_prepare(err_nb), cbad);
      if (cbad) {  // some check failed: the same operations the slow way (bit-identical whenever the checks pass)
#pragma unroll
        for (int d = 0; d < D; d++) {
          const double ay = fabs(y[d]), at = fabs(ytrial[d]);
          const bool ygt = (unsigned long long)__double_as_longlong(ay) > (unsigned long long)__double_as_longlong(at);
          yerr[d] = ynum[d] / (A.abstol + (ygt ? ay : at) * A.reltol);
        }
        err_nb = norm2_small<D>(yerr);
        inv_err = 1.0 / err_nb;
      }
#pragma unroll
      for (int d = 0; d < D; d++) {
        const unsigned int he = (unsigned int)__double2hiint(yerr[d]) & 0x7ff00000u;
        expo_or = expo_or > he ? expo_or : he;
      }
    }
    if (!LITERAL) weird = (expo_or == 0x7ff00000u) || rhs_bad;
    if (!LITERAL && weird) {  // hand the trajectory to the literal pass (see header)
      A.status[idx] = ODE_B200_ST_REDO;
      atomicAdd(A.redo_count, 1);
      idx = -1;
      continue;
    }
    double err, newdt;
    if (nanout) {  // :432
      err = 10.0;
      newdt = dt * 0.2;
      timeout = 5;
    } else {
      err = CNB ? err_nb : norm2_small<D>(yerr);  // :435
      // :436  min(maxstep, tdir*dt*max(facmin, fac*(1/err)^(1/(order+1)))); tdir = +-1 is applied as a
      // sign (exact), min/max keep Julia's NaN propagation (maxstep, facmin and dt are never NaN here)
      const double g = 0.8 * dm_pow(CNB ? inv_err : 1.0 / err, EXPO);
      // max(facmin, g) with g >= +0 or NaN: unsigned compare of the bit patterns (NaN sorts above every number and is
      // propagated, like Julia's max), again off the FP64 pipe
      const double gmax = ((unsigned long long)__double_as_longlong(g) > 0x3FC999999999999AULL) ? g : 0.2;
      double nd = jl_min_nn(A.maxstep, (fwd ? dt : -dt) * gmax);
      const double ndt = jl_min_nn(dt, nd);  // :437-440 (signed dt), branch-free
      nd = (timeout > 0) ? ndt : nd;
      timeout = (timeout > 0) ? timeout - 1 : 0;
      newdt = fwd ? nd : -nd;  // :441
    }
    // :312  err <= 1 for err >= +0 or NaN, as an unsigned compare of bit patterns (NaN compares false either way)
    const bool accept = (unsigned long long)__double_as_longlong(err) <= 0x3FF0000000000000ULL;
    if (MODE == 1 && idx == A.trace_traj) {
      if (ntrace < A.trace_cap) {
        A.trace[ntrace * 4 + 0] = t;
        A.trace[ntrace * 4 + 1] = dt;
        A.trace[ntrace * 4 + 2] = err;
        A.trace[ntrace * 4 + 3] = accept ? 1.0 : 0.0;
      }
      ntrace++;
    }

    int fin = -1;  // >= 0: trajectory ends with this status
    bool fin_trial = false;
    if (accept) {
      nacc++;
      double f1[D];
      if (FSAL) {  // :320
#pragma unroll
        for (int d = 0; d < D; d++) f1[d] = klast[d];
      } else {
        Rhs::eval(t + dt, ytrial, f1, p);  // (exact: this branch is taken once per accepted step)
      }
      if (MODE == 1) {
        double* py = A.pts_y + (idx * A.pts_cap) * D;
        double* pt = A.pts_t + idx * A.pts_cap;
        if (A.points == ODE_B200_POINTS_SPECIFIED) {  // :321-326
          while (iter - 1 < nfixed && (tdir * A.tspan[iter - 1] < tdir * (t + dt) || islast)) {
            if (iter - 1 < A.pts_cap) hermite_point<D>(py + (long long)(iter - 1) * D, A.tspan[iter - 1], t, dt, y, ytrial, k1, f1);
            iter++;
          }
        } else if (A.points == ODE_B200_POINTS_ALL) {  // :327-340
          while (iter_fixed - 1 < nfixed && tdir * t < tdir * A.tspan[iter_fixed - 1] &&
                 tdir * A.tspan[iter_fixed - 1] < tdir * (t + dt)) {
            if (npts < A.pts_cap) {
              hermite_point<D>(py + (long long)npts * D, A.tspan[iter_fixed - 1], t, dt, y, ytrial, k1, f1);
              pt[npts] = A.tspan[iter_fixed - 1];
            }
            npts++;
            iter_fixed++;
          }
          if (npts < A.pts_cap) {
#pragma unroll
            for (int d = 0; d < D; d++) py[(long long)npts * D + d] = ytrial[d];
            pt[npts] = t + dt;
          }
          npts++;
        }
      }
#pragma unroll
      for (int d = 0; d < D; d++) k1[d] = f1[d];  // :341
      if (islast) {                               // :344
        fin = ODE_B200_ST_OK;
        fin_trial = true;
      } else {
#pragma unroll
        for (int d = 0; d < D; d++) y[d] = ytrial[d];  // :347
        t += dt;                                       // :350
        dt = newdt;                                    // :351
        // :354-357  tdir*(t+dt+dt/100) >= tdir*tend, as a signed compare.  The division is skipped when the
        // test cannot pass: dt/100 <= |dt|/16 and rounding is monotone, so (t+dt)+dt/100 <= (t+dt)+|dt|/16
        // (mirror image backwards); a NaN dt fails the filter and takes the exact test.
        const double sdt = t + dt, mdt = fabs(dt) * 0.0625;
        if (!(fwd ? (sdt + mdt < tend) : (sdt - mdt > tend))) {
          const double tl = sdt + dt / 100;
          if (fwd ? (tl >= tend) : (tl <= tend)) {
            dt = tend - t;
            islast = true;
          }
        }
      }
    } else if (fabs(newdt) < A.minstep) {  // :358-360
      fin = ODE_B200_ST_MINSTEP;
    } else {  // :361-366
      islast = false;
      nrej++;
      dt = newdt;
      timeout = 5;
      if (!is_finite(dt)) fin = ODE_B200_ST_NONFINITE;  // the reference would spin forever
    }
    if (fin < 0 && nacc + nrej >= A.max_steps) fin = ODE_B200_ST_MAXSTEPS;  // every attempt so far was counted

    if (fin >= 0) {
      // RHS calls: 2 in hinit, S-1 per attempt (a minstep stop is an attempt in neither counter), +1 per
      // accepted step without FSAL
      const int nrhs = 2 + (S - 1) * (nacc + nrej + (fin == ODE_B200_ST_MINSTEP ? 1 : 0)) + (FSAL ? 0 : nacc);
      if (MODE == 1 && npts > A.pts_cap && A.pts_y) fin = ODE_B200_ST_OVERFLOW;
      if (A.t_final) A.t_final[idx] = fin_trial ? t + dt : t;
      if (A.y_final) {
#pragma unroll
        for (int d = 0; d < D; d++) A.y_final[idx * D + d] = fin_trial ? ytrial[d] : y[d];
      }
      if (A.n_accept) A.n_accept[idx] = nacc;
      if (A.n_reject) A.n_reject[idx] = nrej;
      if (A.n_rhs) A.n_rhs[idx] = nrhs;
      if (A.status) A.status[idx] = fin;
      if (A.status_out) A.status_out[idx] = fin;
      if (A.chunk_done != nullptr) {  // progressive copy-back: the last finisher of a chunk tells the host
        __threadfence();
        const long long c = idx >> A.chunk_log2;
        const long long c_lo = c << A.chunk_log2, c_hi = (c_lo + (1LL << A.chunk_log2) < A.n_traj) ? c_lo + (1LL << A.chunk_log2) : A.n_traj;
        if (atomicAdd(&A.chunk_done[c], 1) == (int)(c_hi - c_lo) - 1) {
          __threadfence_system();
          *(volatile int*)&A.chunk_flag[c] = 1;
        }
      }
      if (MODE == 1) {
        if (A.pts_n) A.pts_n[idx] = npts;
        if (idx == A.trace_traj && A.trace_n) *A.trace_n = ntrace;
      }
      idx = -1;
    }
    if (DRAIN && A.dump != nullptr && idx >= 0 && *(volatile int*)&warp_out[threadIdx.x >> 5] != 0) {
      double* rec = A.dump + (long long)atomicAdd(A.dump_count, 1u) * dump_width(D);
      rec[0] = fabs(t - tstart);
      rec[1] = t;
      rec[2] = dt;
      rec[3] = __longlong_as_double(idx);
      rec[4] = __longlong_as_double((long long)((unsigned long long)(unsigned int)nacc | ((unsigned long long)(unsigned int)nrej << 32)));
      rec[5] = __longlong_as_double((long long)((unsigned long long)(unsigned int)timeout | (islast ? 0x100ULL : 0ULL)));
#pragma unroll
      for (int d = 0; d < D; d++) {
        rec[6 + d] = y[d];
        rec[6 + D + d] = k1[d];
      }
      break;  // (the counter is exhausted: nothing else to fetch)
    }
  }
}

}  // namespace odeb200

// ensemble_fixed.cuh -- fixed-step explicit Runge-Kutta on the tspan grid.
//
// Replaces n_traj calls of oderk_fixed (/root/reference src/runge_kutta.jl:176-212)
// behind ode1 / ode2_midpoint / ode2_heun / ode4 (:165-168).  Every stage,
// including the first, calls F (calc_next_k! with s = 1, :205); the update is
// ys[i+1][d] += (dt*b[s])*ks[s][d] (:207).
#pragma once

#include "common.cuh"
#include "rhs.cuh"
#include "tableaus.cuh"

namespace odeb200 {

template <class Tab, class Rhs>
__global__ void __launch_bounds__(ENS_BLOCK) ens_fixed_kernel(const EnsembleArgs A) {
  constexpr int D = Rhs::DOF, S = Tab::S, NPA = RhsNP<Rhs>::value;
  const int nt = A.n_tspan;
  while (true) {
    long long idx = (long long)atomicAdd(A.work_counter, 1ULL);
    if (idx >= A.n_traj) break;
    double y[D], p[NPA];
#pragma unroll
    for (int d = 0; d < D; d++) y[d] = A.y0[idx * D + d];
#pragma unroll
    for (int q = 0; q < NPA; q++) p[q] = (q < Rhs::NP) ? A.params[idx * A.params_stride + q] : 0.0;
    double* py = A.pts_y ? A.pts_y + (idx * A.pts_cap) * D : nullptr;
    double* pt = A.pts_t ? A.pts_t + idx * A.pts_cap : nullptr;
    if (py && A.pts_cap > 0) {
#pragma unroll
      for (int d = 0; d < D; d++) py[d] = y[d];  // ys[1] = deepcopy(y0) :193
    }
    if (pt && A.pts_cap > 0) pt[0] = A.tspan[0];
    for (int i = 0; i + 1 < nt; i++) {  // :201
      const double ti = A.tspan[i];
      const double dt = A.tspan[i + 1] - ti;  // :202
      double ynew[D], dtk[S][D];
#pragma unroll
      for (int d = 0; d < D; d++) ynew[d] = y[d];  // :203
      static_for<S>([&](auto I) {                  // :204
        constexpr int s = decltype(I)::value;
        double ytmp[D], ks[D];
#pragma unroll
        for (int d = 0; d < D; d++) ytmp[d] = y[d];
        static_for<s>([&](auto J) {
          constexpr int ss = decltype(J)::value;
          constexpr double a = Tab::a(s, ss);
          // every coefficient is multiplied, zeros included, exactly like the reference (Inf*0 = NaN);
          // these kernels are not FP64-critical, so no second "literal" pass is needed here
#pragma unroll
          for (int d = 0; d < D; d++) ytmp[d] = ytmp[d] + dtk[ss][d] * a;  // (dt*k)*a :454
        });
        constexpr double c = Tab::c(s);
        Rhs::eval(ti + c * dt, ytmp, ks, p);  // :456
        constexpr double b = Tab::b(0, s);
#pragma unroll
        for (int d = 0; d < D; d++) {
          dtk[s][d] = dt * ks[d];
          ynew[d] = ynew[d] + (dt * b) * ks[d];  // :207
        }
      });
#pragma unroll
      for (int d = 0; d < D; d++) y[d] = ynew[d];
      if (py && i + 1 < A.pts_cap) {
#pragma unroll
        for (int d = 0; d < D; d++) py[(long long)(i + 1) * D + d] = y[d];
      }
      if (pt && i + 1 < A.pts_cap) pt[i + 1] = A.tspan[i + 1];
    }
    if (A.t_final) A.t_final[idx] = A.tspan[nt - 1];
    if (A.y_final) {
#pragma unroll
      for (int d = 0; d < D; d++) A.y_final[idx * D + d] = y[d];
    }
    if (A.n_accept) A.n_accept[idx] = nt - 1;
    if (A.n_reject) A.n_reject[idx] = 0;
    if (A.n_rhs) A.n_rhs[idx] = (nt - 1) * S;
    if (A.status) A.status[idx] = (A.pts_y && nt > A.pts_cap) ? ODE_B200_ST_OVERFLOW : ODE_B200_ST_OK;
    if (A.pts_n) A.pts_n[idx] = nt;
  }
}

}  // namespace odeb200

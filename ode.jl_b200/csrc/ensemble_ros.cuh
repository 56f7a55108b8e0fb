// ensemble_ros.cuh -- the remaining fixed-step solver families on the tspan grid:
//   * oderosenbrock (/root/reference src/ODE.jl:366-404) with the Shampine (ode4s = ode4s_s,
//     :422-438) and Kaps-Rentrop (ode4s_kr, :408-419) coefficients: per step one Jacobian
//     (fdjacobian or the registered analytic functor), jac = I/(gamma*h) - dF/dx factorised once
//     as the in-register LU of ensemble_ode23s.cuh (the reference's four `jac \ v` re-factorise
//     the same matrix), four RHS calls and four solves;
//   * ode_ms with order 4, Adams-Bashforth (ode4ms, src/ODE.jl:175-212, coefficients :440-443),
//     the last four derivatives kept in registers.
// One trajectory per thread, same work-fetch scheme as the other ensemble kernels.
#pragma once

#include "common.cuh"
#include "ensemble_ode23s.cuh"
#include "rhs.cuh"

namespace odeb200 {

struct RosShampine {  // s4_coefficients, src/ODE.jl:422-431
  static constexpr double gamma = 0.5;
  static constexpr double a[4][4] = {{0, 0, 0, 0}, {2, 0, 0, 0}, {48.0 / 25, 6.0 / 25, 0, 0}, {48.0 / 25, 6.0 / 25, 0, 0}};
  static constexpr double b[4] = {19.0 / 9, 1.0 / 2, 25.0 / 108, 125.0 / 108};
  static constexpr double c[4][4] = {{0, 0, 0, 0}, {-8, 0, 0, 0}, {372.0 / 25, 12.0 / 5, 0, 0}, {-112.0 / 125, -54.0 / 125, -2.0 / 5, 0}};
};
struct RosKapsRentrop {  // kr4_coefficients, src/ODE.jl:408-417
  static constexpr double gamma = 0.231;
  static constexpr double a[4][4] = {{0, 0, 0, 0}, {2, 0, 0, 0}, {4.452470820736, 4.16352878860, 0, 0}, {4.452470820736, 4.16352878860, 0, 0}};
  static constexpr double b[4] = {3.95750374663, 4.62489238836, 0.617477263873, 1.28261294568};
  static constexpr double c[4][4] = {{0, 0, 0, 0}, {-5.07167533877, 0, 0, 0}, {6.02015272865, 0.1597500684673, 0, 0},
                                     {-1.856343618677, -8.50538085819, -2.08407513602, 0}};
};

template <class Rhs, class Coef, int JAC>
__global__ void __launch_bounds__(ENS_BLOCK) ens_rosenbrock_kernel(const EnsembleArgs A) {
  constexpr int D = Rhs::DOF, NPA = RhsNP<Rhs>::value;
  const int nt = A.n_tspan;
  while (true) {
    long long idx = (long long)atomicAdd(A.work_counter, 1ULL);
    if (idx >= A.n_traj) break;
    double x[D], p[NPA];
#pragma unroll
    for (int d = 0; d < D; d++) x[d] = A.y0[idx * D + d];
#pragma unroll
    for (int q = 0; q < NPA; q++) p[q] = (q < Rhs::NP) ? A.params[idx * A.params_stride + q] : 0.0;
    double* py = A.pts_y ? A.pts_y + (idx * A.pts_cap) * D : nullptr;
    double* pt = A.pts_t ? A.pts_t + idx * A.pts_cap : nullptr;
    if (py && A.pts_cap > 0) {
#pragma unroll
      for (int d = 0; d < D; d++) py[d] = x[d];
    }
    if (pt && A.pts_cap > 0) pt[0] = A.tspan[0];
    int nrhs = 0, status = ODE_B200_ST_OK, done = 0;
    for (int st = 0; st + 1 < nt; st++) {  // :379
      const double ts = A.tspan[st], hs = A.tspan[st + 1] - ts;
      double J[D][D];
      nrhs += eval_jacobian<Rhs, JAC, D, NPA>(ts, x, p, J);  // :383
      LuReg<D> W;                                            // :385  jac = I/(gamma*hs) - dFdx
      const double lam = 1.0 / (Coef::gamma * hs);
#pragma unroll
      for (int i = 0; i < D; i++) {
#pragma unroll
        for (int j = 0; j < D; j++) W.a[i][j] = (i == j) ? (-J[i][j] + lam) : -J[i][j];
      }
      W.factor();
      if (W.singular) {  // Julia raises SingularException; report a truncated solution
        status = ODE_B200_ST_SINGULAR;
        break;
      }
      double g[4][D], xn[D], f[D];
      const Recip Rhs_ = recip_prepare(hs);  // dF/hs for three stages x D components
      constexpr double b0 = Coef::b[0];
      Rhs::eval(ts + b0 * hs, x, f, p);  // :388
#pragma unroll
      for (int d = 0; d < D; d++) g[0][d] = f[d];
      W.solve(g[0]);
#pragma unroll
      for (int d = 0; d < D; d++) xn[d] = x[d] + b0 * g[0][d];  // :389
      static_for<3>([&](auto I) {                               // :391
        constexpr int i = decltype(I)::value + 1;
        double dx[D], dF[D], arg[D];
#pragma unroll
        for (int d = 0; d < D; d++) {
          dx[d] = 0.0;
          dF[d] = 0.0;
        }
        static_for<i>([&](auto Jx) {  // :394-397 (every j < i, zero coefficients included, as the reference)
          constexpr int j = decltype(Jx)::value;
          constexpr double aij = Coef::a[i][j], cij = Coef::c[i][j];
#pragma unroll
          for (int d = 0; d < D; d++) {
            dx[d] = dx[d] + aij * g[j][d];
            dF[d] = dF[d] + cij * g[j][d];
          }
        });
#pragma unroll
        for (int d = 0; d < D; d++) arg[d] = x[d] + dx[d];
        constexpr double bi = Coef::b[i];
        Rhs::eval(ts + bi * hs, arg, f, p);  // :398
#pragma unroll
        for (int d = 0; d < D; d++) g[i][d] = f[d] + div_by<true>(dF[d], Rhs_);
        W.solve(g[i]);
#pragma unroll
        for (int d = 0; d < D; d++) xn[d] = xn[d] + bi * g[i][d];  // :399
      });
      nrhs += 4;
#pragma unroll
      for (int d = 0; d < D; d++) x[d] = xn[d];
      done = st + 1;
      if (py && st + 1 < A.pts_cap) {
#pragma unroll
        for (int d = 0; d < D; d++) py[(long long)(st + 1) * D + d] = x[d];
      }
      if (pt && st + 1 < A.pts_cap) pt[st + 1] = A.tspan[st + 1];
    }
    if (status == ODE_B200_ST_OK && A.pts_y && done + 1 > A.pts_cap) status = ODE_B200_ST_OVERFLOW;
    if (A.t_final) A.t_final[idx] = A.tspan[done];
    if (A.y_final) {
#pragma unroll
      for (int d = 0; d < D; d++) A.y_final[idx * D + d] = x[d];
    }
    if (A.n_accept) A.n_accept[idx] = done;
    if (A.n_reject) A.n_reject[idx] = 0;
    if (A.n_rhs) A.n_rhs[idx] = nrhs;
    if (A.status) A.status[idx] = status;
    if (A.pts_n) A.pts_n[idx] = done + 1;
  }
}

// ode4ms: x[i+1] = x[i] + sum_j (h*b[so][j]) * xdot[i-(so-1)+j], so = min(i, 4) (1-based i), :198-207
template <class Rhs>
__global__ void __launch_bounds__(ENS_BLOCK) ens_ms_kernel(const __grid_constant__ EnsembleArgs A) {
  constexpr int D = Rhs::DOF, NPA = RhsNP<Rhs>::value, MO = ODE_B200_MS_MAX_ORDER;
  // b[so-1][j] = A.tab[(so-1)*MO + j]: rows 1-4 are ms_coefficients4 (:440-443), row 5 is filled by the host the way
  // the reference computes it (:183-194)
  const int nt = A.n_tspan, order = A.ms_order;
  while (true) {
    long long idx = (long long)atomicAdd(A.work_counter, 1ULL);
    if (idx >= A.n_traj) break;
    double x[D], p[NPA], hist[MO][D];  // hist[MO-1] = newest derivative
#pragma unroll
    for (int d = 0; d < D; d++) {
      x[d] = A.y0[idx * D + d];
#pragma unroll
      for (int q = 0; q < MO; q++) hist[q][d] = 0.0;
    }
#pragma unroll
    for (int q = 0; q < NPA; q++) p[q] = (q < Rhs::NP) ? A.params[idx * A.params_stride + q] : 0.0;
    double* py = A.pts_y ? A.pts_y + (idx * A.pts_cap) * D : nullptr;
    double* pt = A.pts_t ? A.pts_t + idx * A.pts_cap : nullptr;
    if (py && A.pts_cap > 0) {
#pragma unroll
      for (int d = 0; d < D; d++) py[d] = x[d];
    }
    if (pt && A.pts_cap > 0) pt[0] = A.tspan[0];
    for (int i = 0; i + 1 < nt; i++) {
      const double ti = A.tspan[i], h = A.tspan[i + 1] - ti;
#pragma unroll
      for (int d = 0; d < D; d++) {
#pragma unroll
        for (int q = 0; q + 1 < MO; q++) hist[q][d] = hist[q + 1][d];
      }
      Rhs::eval(ti, x, hist[MO - 1], p);  // :201
      const int so = (i + 1 < order) ? i + 1 : order;  // :200  steporder = min(i, order)
      double hb[MO];  // h[i]*b[steporder, j] for the derivative in hist[q]: j = q - (MO - so), oldest first
#pragma unroll
      for (int q = 0; q < MO; q++) {
        const int j = q - (MO - so);
        hb[q] = (j >= 0) ? h * A.tab[(so - 1) * MO + j] : 0.0;
      }
#pragma unroll
      for (int d = 0; d < D; d++) {  // :203-206, oldest derivative first
        double v = x[d];
#pragma unroll
        for (int q = 0; q < MO; q++)
          if (q >= MO - so) v = v + hb[q] * hist[q][d];
        x[d] = v;
      }
      if (py && i + 1 < A.pts_cap) {
#pragma unroll
        for (int d = 0; d < D; d++) py[(long long)(i + 1) * D + d] = x[d];
      }
      if (pt && i + 1 < A.pts_cap) pt[i + 1] = A.tspan[i + 1];
    }
    if (A.t_final) A.t_final[idx] = A.tspan[nt - 1];
    if (A.y_final) {
#pragma unroll
      for (int d = 0; d < D; d++) A.y_final[idx * D + d] = x[d];
    }
    if (A.n_accept) A.n_accept[idx] = nt - 1;
    if (A.n_reject) A.n_reject[idx] = 0;
    if (A.n_rhs) A.n_rhs[idx] = nt - 1;
    if (A.status) A.status[idx] = (A.pts_y && nt > A.pts_cap) ? ODE_B200_ST_OVERFLOW : ODE_B200_ST_OK;
    if (A.pts_n) A.pts_n[idx] = nt;
  }
}

}  // namespace odeb200

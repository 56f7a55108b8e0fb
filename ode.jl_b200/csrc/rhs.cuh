// rhs.cuh -- registered right-hand sides F(t, y; p) as compiled device functors.
//
// A Julia closure cannot cross the C ABI, so `F` is a registered id
// (include/ode_b200.h ode_b200_rhs) plus a parameter vector.  Each functor states
// the same expression, in the same association, as the Julia `F` the reference
// would be given; the CPU oracle states them independently
// (oracle/ode_oracle.cpp Rhs::operator(), oracle/ode_oracle.py rhs_factory) and
// the parity tests compare bit for bit.  No FMA contraction (-fmad=false).
// `jac` is the registered analytic Jacobian dF/dy -- what a user would pass as the
// `jacobian = G(t,y)` keyword of ode23s / ode4s (src/ODE.jl:262-263, :368-369).
#pragma once

namespace odeb200 {

// f(u,p,t) = a*u            README.md:19 (a = 1.01), test/runtests.jl:54 (a = 1)
struct RhsLinear {
  static constexpr int DOF = 1, NP = 1, ID = ODE_B200_RHS_LINEAR;
  static constexpr bool AUTONOMOUS = true;  // F does not read t
  __device__ __forceinline__ static void eval(double, const double (&y)[1], double (&f)[1], const double (&p)[1]) {
    f[0] = p[0] * y[0];
  }
  __device__ __forceinline__ static void jac(double, const double (&)[1], double (&J)[1][1], const double (&p)[1]) {
    J[0][0] = p[0];
  }
};
// (t,y) -> T(6)             test/runtests.jl:40
struct RhsConst {
  static constexpr int DOF = 1, NP = 1, ID = ODE_B200_RHS_CONST;
  static constexpr bool AUTONOMOUS = true;  // F does not read t
  __device__ __forceinline__ static void eval(double, const double (&)[1], double (&f)[1], const double (&p)[1]) {
    f[0] = p[0];
  }
  __device__ __forceinline__ static void jac(double, const double (&)[1], double (&J)[1][1], const double (&)[1]) {
    J[0][0] = 0.0;
  }
};
// (t,y) -> T(2)*t           test/runtests.jl:47
struct RhsTimeLin {
  static constexpr int DOF = 1, NP = 1, ID = ODE_B200_RHS_TIMELIN;
  __device__ __forceinline__ static void eval(double t, const double (&)[1], double (&f)[1], const double (&p)[1]) {
    f[0] = p[0] * t;
  }
  __device__ __forceinline__ static void jac(double, const double (&)[1], double (&J)[1][1], const double (&)[1]) {
    J[0][0] = 0.0;
  }
};
// (t,y) -> [-y[2]; y[1]]    test/runtests.jl:67
struct RhsRotation {
  static constexpr int DOF = 2, NP = 0, ID = ODE_B200_RHS_ROTATION;
  static constexpr bool AUTONOMOUS = true;  // F does not read t
  __device__ __forceinline__ static void eval(double, const double (&y)[2], double (&f)[2], const double (&)[1]) {
    f[0] = -y[1];
    f[1] = y[0];
  }
  __device__ __forceinline__ static void jac(double, const double (&)[2], double (&J)[2][2], const double (&)[1]) {
    J[0][0] = 0.0;
    J[0][1] = -1.0;
    J[1][0] = 1.0;
    J[1][1] = 0.0;
  }
};
// Lorenz, examples/Lorenz_Attractor.ipynb cell 4: [s*(y-x); x*(r-z)-y; x*y-b*z]
struct RhsLorenz {
  static constexpr int DOF = 3, NP = 3, ID = ODE_B200_RHS_LORENZ;
  static constexpr bool AUTONOMOUS = true;  // F does not read t
  __device__ __forceinline__ static void eval(double, const double (&y)[3], double (&f)[3], const double (&p)[3]) {
    f[0] = p[0] * (y[1] - y[0]);
    f[1] = y[0] * (p[1] - y[2]) - y[1];
    f[2] = y[0] * y[1] - p[2] * y[2];
  }
  __device__ __forceinline__ static void jac(double, const double (&y)[3], double (&J)[3][3], const double (&p)[3]) {
    J[0][0] = -p[0];
    J[0][1] = p[0];
    J[0][2] = 0.0;
    J[1][0] = p[1] - y[2];
    J[1][1] = -1.0;
    J[1][2] = -y[0];
    J[2][0] = y[1];
    J[2][1] = y[0];
    J[2][2] = -p[2];
  }
};
// Van der Pol: [y2; mu*((1-y1*y1)*y2) - y1]
struct RhsVdp {
  static constexpr int DOF = 2, NP = 1, ID = ODE_B200_RHS_VDP;
  static constexpr bool AUTONOMOUS = true;  // F does not read t
  __device__ __forceinline__ static void eval(double, const double (&y)[2], double (&f)[2], const double (&p)[1]) {
    f[0] = y[1];
    f[1] = p[0] * ((1.0 - y[0] * y[0]) * y[1]) - y[0];
  }
  __device__ __forceinline__ static void jac(double, const double (&y)[2], double (&J)[2][2], const double (&p)[1]) {
    J[0][0] = 0.0;
    J[0][1] = 1.0;
    J[1][0] = p[0] * ((-2.0 * y[0]) * y[1]) - 1.0;
    J[1][1] = p[0] * (1.0 - y[0] * y[0]);
  }
};
// Robertson, test/runtests.jl:81-87 (ydot[1], then ydot[3], then ydot[2] = -ydot[1]-ydot[3])
struct RhsRobertson {
  static constexpr int DOF = 3, NP = 3, ID = ODE_B200_RHS_ROBERTSON;
  static constexpr bool AUTONOMOUS = true;  // F does not read t
  __device__ __forceinline__ static void eval(double, const double (&y)[3], double (&f)[3], const double (&p)[3]) {
    double d1 = -p[0] * y[0] + p[2] * y[1] * y[2];
    double d3 = p[1] * y[1] * y[1];
    f[0] = d1;
    f[2] = d3;
    f[1] = -d1 - d3;
  }
  __device__ __forceinline__ static void jac(double, const double (&y)[3], double (&J)[3][3], const double (&p)[3]) {
    const double a = p[2] * y[2], b = p[2] * y[1], c = (2.0 * p[1]) * y[1];
    J[0][0] = -p[0];
    J[0][1] = a;
    J[0][2] = b;
    J[1][0] = p[0];
    J[1][1] = -a - c;
    J[1][2] = -b;
    J[2][0] = 0.0;
    J[2][1] = c;
    J[2][2] = 0.0;
  }
};
// planar Kepler, mu = 1: r2 = x*x+y*y; r3 = r2*sqrt(r2); [vx, vy, -x/r3, -y/r3]
struct RhsKepler {
  static constexpr int DOF = 4, NP = 0, ID = ODE_B200_RHS_KEPLER;
  static constexpr bool AUTONOMOUS = true;  // F does not read t
  __device__ __forceinline__ static void eval(double, const double (&y)[4], double (&f)[4], const double (&)[1]) {
    double r2 = y[0] * y[0] + y[1] * y[1];
    double r3 = r2 * sqrt(r2);
    const Recip R3 = recip_prepare(r3);  // two numerators over one denominator (common.cuh)
    f[0] = y[2];
    f[1] = y[3];
    f[2] = div_by(-y[0], R3);
    f[3] = div_by(-y[1], R3);
  }
  // The same values as straight-line code for the fast ensemble kernels: sqrt and the two divisions carry deferred
  // acceptance checks (common.cuh sqrt_nb / div_nb) instead of a branch each; `bad` sends the trajectory to the
  // literal pass, which uses eval().
  __device__ __forceinline__ static void eval_nb(double, const double (&y)[4], double (&f)[4], const double (&)[1], bool& bad) {
    double r2 = y[0] * y[0] + y[1] * y[1];
    double r3 = r2 * sqrt_nb(r2, bad);
    const Recip R3 = recip_prepare(r3);
    f[0] = y[2];
    f[1] = y[3];
    f[2] = div_nb<true>(-y[0], R3, bad);
    f[3] = div_nb<true>(-y[1], R3, bad);
  }
  __device__ __forceinline__ static void jac(double, const double (&y)[4], double (&J)[4][4], const double (&)[1]) {
    const double x = y[0], yy = y[1];
    const double r2 = x * x + yy * yy;
    const double r3 = r2 * sqrt(r2), r5 = r3 * r2;
#pragma unroll
    for (int i = 0; i < 4; i++) {
#pragma unroll
      for (int j = 0; j < 4; j++) J[i][j] = 0.0;
    }
    J[0][2] = 1.0;
    J[1][3] = 1.0;
    J[2][0] = (3.0 * x) * x / r5 - 1.0 / r3;
    J[2][1] = (3.0 * x) * yy / r5;
    J[3][0] = (3.0 * x) * yy / r5;
    J[3][1] = (3.0 * yy) * yy / r5 - 1.0 / r3;
  }
};

// 1-D reaction-diffusion stencil (large-system regime): point value from the three neighbours,
// kappa*((um - 2*u) + up) + (r*u)*(1 - u).  Stencil functors see a window w[k] = u(i-R+k), the parameter
// vector, the stage time and the global grid index (csrc/large.cuh).
struct RhsReactionDiffusion {
  static constexpr int NP = 2, RADIUS = 1, ID = ODE_B200_RHS_REACTION_DIFFUSION;
  __device__ __forceinline__ static double point(const double (&w)[3], const double (&p)[8], double, long long, long long) {
    return p[0] * ((w[0] - 2.0 * w[1]) + w[2]) + (p[1] * w[1]) * (1.0 - w[1]);
  }
};

// R::AUTONOMOUS (optional member): the functor ignores its time argument, so F(t1, y) and F(t2, y) are the same
// bits; absent = false (run-time registered functors unless they declare it).
template <class R, class = void>
struct RhsAutonomous {
  static constexpr bool value = false;
};
template <class R>
struct RhsAutonomous<R, decltype((void)R::AUTONOMOUS)> {
  static constexpr bool value = R::AUTONOMOUS;
};

template <class R>
struct RhsNP {
  static constexpr int value = R::NP > 0 ? R::NP : 1;
};

// rhs_eval_fast: R::eval_nb (deferred checks ORed into `bad`) where a functor has one, else R::eval
template <class R, int D, int NPA>
__device__ __forceinline__ auto rhs_eval_fast_impl(double t, const double (&y)[D], double (&f)[D], const double (&p)[NPA],
                                                   bool& bad, int) -> decltype(R::eval_nb(t, y, f, p, bad), void()) {
  R::eval_nb(t, y, f, p, bad);
}
template <class R, int D, int NPA>
__device__ __forceinline__ void rhs_eval_fast_impl(double t, const double (&y)[D], double (&f)[D], const double (&p)[NPA],
                                                   bool&, long) {
  R::eval(t, y, f, p);
}
#ifndef ODEB200_RHS_NB
// 1: functors with an eval_nb use it in the fast adaptive kernels; 0: always R::eval (a branch per sqrt / division).
// Measured on Kepler/feh78 (profiles/ensemble_kernels_r2.md): with the branches gone the scheduler interleaves so
// much more that the kernel needs 214 registers instead of 150 (2 CTAs/SM): 22.3 ms vs 18.6 ms; capped at 168 / 128
// registers it spills: 19.2 / 19.0 ms.  Off.
#define ODEB200_RHS_NB 0
#endif
template <class R, int D, int NPA>
__device__ __forceinline__ void rhs_eval_fast(double t, const double (&y)[D], double (&f)[D], const double (&p)[NPA], bool& bad) {
#if ODEB200_RHS_NB
  rhs_eval_fast_impl<R, D, NPA>(t, y, f, p, bad, 0);
#else
  R::eval(t, y, f, p);
#endif
}

}  // namespace odeb200

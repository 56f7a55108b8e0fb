// dispatch.cuh -- (method, rhs, mode) -> kernel pointer.  One translation unit per
// tableau keeps nvcc compile time parallel (see Makefile).
#pragma once
#include "common.cuh"

namespace odeb200 {

// mode: 0 = final state + counters only, 1 = tout/yout points + trace
// literal: 0 = fast pass (zero coefficients dropped, marks non-finite trajectories), 1 = literal redo pass,
//          2 = fast pass, mode 0, one parameter vector for the whole ensemble (read from the constant bank)
EnsembleKernel adapt_kernel_rk21(int rhs, int mode, int literal);
EnsembleKernel adapt_kernel_rk23(int rhs, int mode, int literal);
EnsembleKernel adapt_kernel_rk45(int rhs, int mode, int literal);
EnsembleKernel adapt_kernel_dopri5(int rhs, int mode, int literal);
EnsembleKernel adapt_kernel_feh78(int rhs, int mode, int literal);
EnsembleKernel adapt_kernel_ck45(int rhs, int mode, int literal);
EnsembleKernel adapt_kernel_rk21_init(int rhs);
const void* adapt_kernel_rk21_stream_init(int rhs);  // hinit pre-pass (ens_adapt_init_kernel)
EnsembleKernel adapt_kernel_rk23_init(int rhs);
const void* adapt_kernel_rk23_stream_init(int rhs);  // hinit pre-pass (ens_adapt_init_kernel)
EnsembleKernel adapt_kernel_rk45_init(int rhs);
const void* adapt_kernel_rk45_stream_init(int rhs);  // hinit pre-pass (ens_adapt_init_kernel)
EnsembleKernel adapt_kernel_dopri5_init(int rhs);
const void* adapt_kernel_dopri5_stream_init(int rhs);  // hinit pre-pass (ens_adapt_init_kernel)
EnsembleKernel adapt_kernel_feh78_init(int rhs);
const void* adapt_kernel_feh78_stream_init(int rhs);  // hinit pre-pass (ens_adapt_init_kernel)
EnsembleKernel adapt_kernel_ck45_init(int rhs);
const void* adapt_kernel_ck45_stream_init(int rhs);   // hinit pre-pass (ens_adapt_init_kernel)
EnsembleKernel fixed_kernel(int method, int rhs);
EnsembleKernel ode23s_kernel(int rhs, int mode, int jac);  // jac: 0 fdjacobian, 1 analytic functor
EnsembleKernel ode23s_init_kernel(int rhs, int jac);       // pre-pass: first step, F0, first Jacobian
EnsembleKernel rosenbrock_kernel(int method, int rhs, int jac);
EnsembleKernel ms_kernel(int rhs);  // ode_ms, orders 1..5

#define ODEB200_RHS_SWITCH(rhs, X)                 \
  switch (rhs) {                                   \
    case ODE_B200_RHS_LINEAR: X(RhsLinear);        \
    case ODE_B200_RHS_CONST: X(RhsConst);          \
    case ODE_B200_RHS_TIMELIN: X(RhsTimeLin);      \
    case ODE_B200_RHS_ROTATION: X(RhsRotation);    \
    case ODE_B200_RHS_LORENZ: X(RhsLorenz);        \
    case ODE_B200_RHS_VDP: X(RhsVdp);              \
    case ODE_B200_RHS_ROBERTSON: X(RhsRobertson);  \
    case ODE_B200_RHS_KEPLER: X(RhsKepler);        \
    default: return nullptr;                       \
  }

#define ODEB200_DEFINE_ADAPT(NAME, TAB)                                                        \
  EnsembleKernel NAME(int rhs, int mode, int literal) {                                        \
    if (literal == 2) {                                                                        \
      ODEB200_RHS_SWITCH(rhs, ODEB200_ADAPT_M0U_##TAB)                                         \
    }                                                                                          \
    if (literal) {                                                                             \
      if (mode) {                                                                              \
        ODEB200_RHS_SWITCH(rhs, ODEB200_ADAPT_M1L_##TAB)                                       \
      } else {                                                                                 \
        ODEB200_RHS_SWITCH(rhs, ODEB200_ADAPT_M0L_##TAB)                                       \
      }                                                                                        \
    }                                                                                          \
    if (mode) {                                                                                \
      ODEB200_RHS_SWITCH(rhs, ODEB200_ADAPT_M1_##TAB)                                          \
    } else {                                                                                   \
      ODEB200_RHS_SWITCH(rhs, ODEB200_ADAPT_M0_##TAB)                                          \
    }                                                                                          \
  }                                                                                            \
  EnsembleKernel NAME##_init(int rhs) { ODEB200_RHS_SWITCH(rhs, ODEB200_ADAPT_INIT_##TAB) }                      \
  const void* NAME##_stream_init(int rhs) { ODEB200_RHS_SWITCH(rhs, ODEB200_ADAPT_SINIT_##TAB) }

}  // namespace odeb200

// ode23s ensemble kernels, plus the fixed-step Rosenbrock (ode4s) and Adams-Bashforth (ode4ms) families.
#include "dispatch.cuh"
#include "ensemble_ode23s.cuh"
#include "ensemble_ros.cuh"
namespace odeb200 {
template <int MODE, int JAC>
static EnsembleKernel ode23s_for(int rhs) {
#define X(R) return ens_ode23s_kernel<R, MODE, JAC>
  ODEB200_RHS_SWITCH(rhs, X)
#undef X
}
EnsembleKernel ode23s_kernel(int rhs, int mode, int jac) {
  if (mode) return jac ? ode23s_for<1, 1>(rhs) : ode23s_for<1, 0>(rhs);
  return jac ? ode23s_for<0, 1>(rhs) : ode23s_for<0, 0>(rhs);
}
template <int JAC>
static EnsembleKernel ode23s_init_for(int rhs) {
#define X(R) return ens_ode23s_init_kernel<R, JAC>
  ODEB200_RHS_SWITCH(rhs, X)
#undef X
}
EnsembleKernel ode23s_init_kernel(int rhs, int jac) { return jac ? ode23s_init_for<1>(rhs) : ode23s_init_for<0>(rhs); }
template <class Coef, int JAC>
static EnsembleKernel ros_for(int rhs) {
#define X(R) return ens_rosenbrock_kernel<R, Coef, JAC>
  ODEB200_RHS_SWITCH(rhs, X)
#undef X
}
EnsembleKernel rosenbrock_kernel(int method, int rhs, int jac) {
  if (method == ODE_B200_M_ODE4S_KR) return jac ? ros_for<RosKapsRentrop, 1>(rhs) : ros_for<RosKapsRentrop, 0>(rhs);
  return jac ? ros_for<RosShampine, 1>(rhs) : ros_for<RosShampine, 0>(rhs);
}
EnsembleKernel ms_kernel(int rhs) {
#define X(R) return ens_ms_kernel<R>
  ODEB200_RHS_SWITCH(rhs, X)
#undef X
}
}  // namespace odeb200

// ode23s ensemble kernels.
#include "dispatch.cuh"
#include "ensemble_ode23s.cuh"
namespace odeb200 {
EnsembleKernel ode23s_kernel(int rhs, int mode) {
  if (mode) {
#define X(R) return ens_ode23s_kernel<R, 1>
    ODEB200_RHS_SWITCH(rhs, X)
#undef X
  } else {
#define X(R) return ens_ode23s_kernel<R, 0>
    ODEB200_RHS_SWITCH(rhs, X)
#undef X
  }
}
}  // namespace odeb200

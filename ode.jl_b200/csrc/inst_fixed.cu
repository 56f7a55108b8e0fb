// Fixed-step ensemble kernels (ode1, ode2_midpoint, ode2_heun, ode4).
#include "dispatch.cuh"
#include "ensemble_fixed.cuh"
namespace odeb200 {
template <class Tab>
static EnsembleKernel fixed_for(int rhs) {
#define X(R) return ens_fixed_kernel<Tab, R>
  ODEB200_RHS_SWITCH(rhs, X)
#undef X
}
EnsembleKernel fixed_kernel(int method, int rhs) {
  switch (method) {
    case ODE_B200_M_FEULER: return fixed_for<TabFEuler>(rhs);
    case ODE_B200_M_MIDPOINT: return fixed_for<TabMidpoint>(rhs);
    case ODE_B200_M_HEUN: return fixed_for<TabHeun>(rhs);
    case ODE_B200_M_RK4: return fixed_for<TabRK4>(rhs);
    default: return nullptr;
  }
}
}  // namespace odeb200

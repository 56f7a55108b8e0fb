// large.cu -- single large method-of-lines system (placeholder until the fused
// stage kernels land; every entry point reports ODE_B200_E_UNSUPPORTED).
#include <cuda_runtime.h>

#include "../../include/ode_b200.h"
#include "host_util.cuh"

using namespace odeb200;

extern "C" {
int ode_b200_solve_large(const ode_b200_opts*, int64_t, const double*, const double*, double, double, double*,
                         ode_b200_large_stats*, double*, int64_t, int64_t*) {
  return set_err(ODE_B200_E_UNSUPPORTED, "large-system path not built yet");
}
int ode_b200_slab_create(const ode_b200_opts*, int64_t, int64_t, int32_t, int32_t, const double*, double, double, void*,
                         ode_b200_slab** out) {
  if (out) *out = nullptr;
  return set_err(ODE_B200_E_UNSUPPORTED, "large-system path not built yet");
}
int ode_b200_slab_set_state(ode_b200_slab*, const double*) { return set_err(ODE_B200_E_UNSUPPORTED, "not built"); }
int ode_b200_slab_get_state(ode_b200_slab*, double*) { return set_err(ODE_B200_E_UNSUPPORTED, "not built"); }
int ode_b200_slab_init_phase(ode_b200_slab*, int, double*, const double*) { return set_err(ODE_B200_E_UNSUPPORTED, "not built"); }
int ode_b200_slab_attempt(ode_b200_slab*, double*) { return set_err(ODE_B200_E_UNSUPPORTED, "not built"); }
int ode_b200_slab_control(ode_b200_slab*, const double*) { return set_err(ODE_B200_E_UNSUPPORTED, "not built"); }
int ode_b200_slab_poll(ode_b200_slab*, double*) { return set_err(ODE_B200_E_UNSUPPORTED, "not built"); }
void ode_b200_slab_free(ode_b200_slab*) {}
}

// host_util.cuh -- host-side helpers shared by the C ABI translation units.
#pragma once
#include <cuda_runtime.h>

namespace odeb200 {
int set_err(int code, const char* fmt, ...);  // fills the thread-local last_error, returns code
void note_launch();                           // counts kernel launches (ode_b200_launch_count)
int ms_order(int method);                     // order of ode_ms for a method id, 0 if it is not a multistep method
double ms_coefficient(int order, int j);      // b[order][j] of ode_ms, j 0-based (capi.cu)
// the embedded (adaptive) explicit RK tableaus: the reference's five (ids 4..8) and the Cash-Karp extension
inline bool method_is_adaptive_rk(int method) { return (method >= 4 && method <= 8) || method == 17; }
}  // namespace odeb200

#define ODEB200_CUDA(call)                                                                              \
  do {                                                                                                  \
    cudaError_t e__ = (call);                                                                           \
    if (e__ != cudaSuccess) {                                                                           \
      cudaGetLastError();                                                                               \
      return ::odeb200::set_err(ODE_B200_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), \
                                __FILE__, __LINE__);                                                    \
    }                                                                                                   \
  } while (0)

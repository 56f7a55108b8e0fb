// Ensemble kernels for the TabRK45 tableau (all registered small-system RHS).
#include "dispatch.cuh"
#include "ensemble_adapt.cuh"
namespace odeb200 {
#define ODEB200_ADAPT_M0_TabRK45(R) return ens_adapt_kernel<TabRK45, R, 0, false>
#define ODEB200_ADAPT_M1_TabRK45(R) return ens_adapt_kernel<TabRK45, R, 1, false>
#define ODEB200_ADAPT_M0L_TabRK45(R) return ens_adapt_kernel<TabRK45, R, 0, true>
#define ODEB200_ADAPT_M1L_TabRK45(R) return ens_adapt_kernel<TabRK45, R, 1, true>
#define ODEB200_ADAPT_M0U_TabRK45(R) return ens_adapt_kernel<TabRK45, R, 0, false, true>
#define ODEB200_ADAPT_INIT_TabRK45(R) return ens_adapt_init_kernel<TabRK45, R>
#define ODEB200_ADAPT_SINIT_TabRK45(R) return (const void*)ens_adapt_stream_init_kernel<TabRK45, R>
ODEB200_DEFINE_ADAPT(adapt_kernel_rk45, TabRK45)
}  // namespace odeb200

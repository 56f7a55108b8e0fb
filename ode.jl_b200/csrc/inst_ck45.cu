// Ensemble kernels for the TabCK45 tableau (all registered small-system RHS).
#include "dispatch.cuh"
#include "ensemble_adapt.cuh"
namespace odeb200 {
#define ODEB200_ADAPT_M0_TabCK45(R) return ens_adapt_kernel<TabCK45, R, 0, false>
#define ODEB200_ADAPT_M1_TabCK45(R) return ens_adapt_kernel<TabCK45, R, 1, false>
#define ODEB200_ADAPT_M0L_TabCK45(R) return ens_adapt_kernel<TabCK45, R, 0, true>
#define ODEB200_ADAPT_M1L_TabCK45(R) return ens_adapt_kernel<TabCK45, R, 1, true>
#define ODEB200_ADAPT_M0U_TabCK45(R) return ens_adapt_kernel<TabCK45, R, 0, false, true>
#define ODEB200_ADAPT_INIT_TabCK45(R) return ens_adapt_init_kernel<TabCK45, R>
#define ODEB200_ADAPT_SINIT_TabCK45(R) return (const void*)ens_adapt_stream_init_kernel<TabCK45, R>
ODEB200_DEFINE_ADAPT(adapt_kernel_ck45, TabCK45)
}  // namespace odeb200

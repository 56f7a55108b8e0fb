// Ensemble kernels for the TabFeh78 tableau (all registered small-system RHS).
#include "dispatch.cuh"
#include "ensemble_adapt.cuh"
namespace odeb200 {
#define ODEB200_ADAPT_M0_TabFeh78(R) return ens_adapt_kernel<TabFeh78, R, 0, false>
#define ODEB200_ADAPT_M1_TabFeh78(R) return ens_adapt_kernel<TabFeh78, R, 1, false>
#define ODEB200_ADAPT_M0L_TabFeh78(R) return ens_adapt_kernel<TabFeh78, R, 0, true>
#define ODEB200_ADAPT_M1L_TabFeh78(R) return ens_adapt_kernel<TabFeh78, R, 1, true>
#define ODEB200_ADAPT_M0U_TabFeh78(R) return ens_adapt_kernel<TabFeh78, R, 0, false, true>
#define ODEB200_ADAPT_INIT_TabFeh78(R) return ens_adapt_init_kernel<TabFeh78, R>
#define ODEB200_ADAPT_SINIT_TabFeh78(R) return (const void*)ens_adapt_stream_init_kernel<TabFeh78, R>
ODEB200_DEFINE_ADAPT(adapt_kernel_feh78, TabFeh78)
}  // namespace odeb200

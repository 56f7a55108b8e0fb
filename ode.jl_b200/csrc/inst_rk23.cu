// Ensemble kernels for the TabRK23 tableau (all registered small-system RHS).
#include "dispatch.cuh"
#include "ensemble_adapt.cuh"
namespace odeb200 {
#define ODEB200_ADAPT_M0_TabRK23(R) return ens_adapt_kernel<TabRK23, R, 0, false>
#define ODEB200_ADAPT_M1_TabRK23(R) return ens_adapt_kernel<TabRK23, R, 1, false>
#define ODEB200_ADAPT_M0L_TabRK23(R) return ens_adapt_kernel<TabRK23, R, 0, true>
#define ODEB200_ADAPT_M1L_TabRK23(R) return ens_adapt_kernel<TabRK23, R, 1, true>
#define ODEB200_ADAPT_M0U_TabRK23(R) return ens_adapt_kernel<TabRK23, R, 0, false, true>
#define ODEB200_ADAPT_INIT_TabRK23(R) return ens_adapt_init_kernel<TabRK23, R>
#define ODEB200_ADAPT_SINIT_TabRK23(R) return (const void*)ens_adapt_stream_init_kernel<TabRK23, R>
ODEB200_DEFINE_ADAPT(adapt_kernel_rk23, TabRK23)
}  // namespace odeb200

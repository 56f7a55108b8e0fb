// Ensemble kernels for the TabDopri5 tableau (all registered small-system RHS).
#include "dispatch.cuh"
#include "ensemble_adapt.cuh"
namespace odeb200 {
#define ODEB200_ADAPT_M0_TabDopri5(R) return ens_adapt_kernel<TabDopri5, R, 0, false>
#define ODEB200_ADAPT_M1_TabDopri5(R) return ens_adapt_kernel<TabDopri5, R, 1, false>
#define ODEB200_ADAPT_M0L_TabDopri5(R) return ens_adapt_kernel<TabDopri5, R, 0, true>
#define ODEB200_ADAPT_M1L_TabDopri5(R) return ens_adapt_kernel<TabDopri5, R, 1, true>
#define ODEB200_ADAPT_M0U_TabDopri5(R) return ens_adapt_kernel<TabDopri5, R, 0, false, true>
#define ODEB200_ADAPT_INIT_TabDopri5(R) return ens_adapt_init_kernel<TabDopri5, R>
#define ODEB200_ADAPT_SINIT_TabDopri5(R) return (const void*)ens_adapt_stream_init_kernel<TabDopri5, R>
ODEB200_DEFINE_ADAPT(adapt_kernel_dopri5, TabDopri5)
}  // namespace odeb200

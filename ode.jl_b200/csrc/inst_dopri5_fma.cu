// The opt-in FMA build (opts.fp_mode = ODE_B200_FP_FMA) of the Dormand-Prince ensemble kernels: the very same
// templates as inst_dopri5.cu, compiled with FMA contraction ON (-fmad=true, see the Makefile rule for *_fma.cu).
// NOT the parity path: a*b+c rounds once instead of twice, so results differ from the reference's in the last
// bits and step counts can differ; it exists to show what the design reaches against the FP64 FMA peak, which
// strict (reference) arithmetic cannot use by construction.  Everything lives in its own namespace so that the two
// builds of the same templates do not collide at link time.
#define odeb200 odeb200_fma
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
#undef odeb200
// type-erased getters for capi.cu (EnsembleArgs has the same layout in both namespaces: same header)
extern "C" const void* odeb200_fma_dopri5_kernel(int rhs, int mode, int literal) {
  return (const void*)odeb200_fma::adapt_kernel_dopri5(rhs, mode, literal);
}
extern "C" const void* odeb200_fma_dopri5_init_kernel(int rhs) { return (const void*)odeb200_fma::adapt_kernel_dopri5_init(rhs); }

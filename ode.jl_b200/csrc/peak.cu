// peak.cu -- FP64 issue-rate microbenchmark (roofline denominator for the
// ensemble regime; MEASURED_PEAKS.json has no FP64 entry).
//
// Each thread runs 16 independent register-resident dependency chains, so the
// FP64 pipe is the only limiter.  kind 0: DFMA chains; kind 1: alternating
// DADD / DMUL chains (what the strict no-FMA solver arithmetic can issue).
#include <cuda_runtime.h>

#include "../../include/ode_b200.h"
#include "../../include/ode_b200_detmath.h"
#include "host_util.cuh"

namespace odeb200 {

template <int KIND>
__global__ void __launch_bounds__(256) fp64_peak_kernel(double* out, int iters, double a, double b) {
  double x[16];
#pragma unroll
  for (int i = 0; i < 16; i++) x[i] = (double)(threadIdx.x + i) * 1e-3;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 16; i++) {
      if (KIND == 0) {
        x[i] = __fma_rn(x[i], a, b);
      } else {
        x[i] = __dmul_rn(x[i], a);
        x[i] = __dadd_rn(x[i], b);
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) s += x[i];
  if (s == 123.456) out[0] = s;  // keep the chains alive
}

}  // namespace odeb200

extern "C" double ode_b200_fp64_peak(int device, int kind) {
  using namespace odeb200;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) {
    cudaGetLastError();
    set_err(ODE_B200_E_NO_DEVICE, "no CUDA device %d", device);
    return -1.0;
  }
  if (cudaSetDevice(device) != cudaSuccess) return -1.0;
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  double* out = nullptr;
  if (cudaMalloc(&out, 8) != cudaSuccess) return -1.0;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int iters = 4096, blocks = sms * 8, threads = 256;
  double best = 0.0;
  for (int rep = 0; rep < 5; rep++) {
    cudaEventRecord(e0);
    if (kind == 0)
      fp64_peak_kernel<0><<<blocks, threads>>>(out, iters, 0.999999, 1e-9);
    else
      fp64_peak_kernel<1><<<blocks, threads>>>(out, iters, 0.999999, 1e-9);
    note_launch();
    cudaEventRecord(e1);
    if (cudaEventSynchronize(e1) != cudaSuccess) {
      best = -1.0;
      break;
    }
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    double instr = (double)blocks * threads * iters * 16.0 * (kind == 0 ? 1.0 : 2.0);
    double rate = instr / (ms * 1e-3);
    if (rep > 0 && rate > best) best = rate;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  return best;
}

namespace odeb200 {
__global__ void detmath_kernel(const double* x, const double* y, double* op, double* ol, long long n) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < n) {
    op[i] = dm_pow(x[i], y[i]);
    ol[i] = dm_log10(x[i]);
  }
}
}  // namespace odeb200

extern "C" int ode_b200_selftest_detmath(int device, const double* x, const double* y, double* out_pow,
                                         double* out_log10, int64_t n) {
  using namespace odeb200;
  int cnt = 0;
  if (cudaGetDeviceCount(&cnt) != cudaSuccess || device < 0 || device >= cnt) {
    cudaGetLastError();
    return set_err(ODE_B200_E_NO_DEVICE, "no CUDA device %d", device);
  }
  ODEB200_CUDA(cudaSetDevice(device));
  double *dx, *dy, *dp, *dl;
  ODEB200_CUDA(cudaMalloc(&dx, n * 8));
  ODEB200_CUDA(cudaMalloc(&dy, n * 8));
  ODEB200_CUDA(cudaMalloc(&dp, n * 8));
  ODEB200_CUDA(cudaMalloc(&dl, n * 8));
  ODEB200_CUDA(cudaMemcpy(dx, x, n * 8, cudaMemcpyHostToDevice));
  ODEB200_CUDA(cudaMemcpy(dy, y, n * 8, cudaMemcpyHostToDevice));
  detmath_kernel<<<(unsigned)((n + 255) / 256), 256>>>(dx, dy, dp, dl, n);
  note_launch();
  ODEB200_CUDA(cudaGetLastError());
  ODEB200_CUDA(cudaMemcpy(out_pow, dp, n * 8, cudaMemcpyDeviceToHost));
  ODEB200_CUDA(cudaMemcpy(out_log10, dl, n * 8, cudaMemcpyDeviceToHost));
  cudaFree(dx);
  cudaFree(dy);
  cudaFree(dp);
  cudaFree(dl);
  return 0;
}

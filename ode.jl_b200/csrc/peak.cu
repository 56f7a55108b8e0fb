// peak.cu -- FP64 issue-rate microbenchmark (roofline denominator for the
// ensemble regime; MEASURED_PEAKS.json has no FP64 entry).
//
// Each thread runs 16 independent register-resident dependency chains, so the
// FP64 pipe is the only limiter.  kind 0: DFMA chains; kind 1: alternating
// DADD / DMUL chains (what the strict no-FMA solver arithmetic can issue).
#include <cuda_runtime.h>

#include "../../include/ode_b200.h"
#include "../../include/ode_b200_detmath.h"
#include "common.cuh"
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
// div_by(a, recip_prepare(b)) against the compiler's a / b on pseudo-random bit patterns: uniform 64-bit
// patterns (all exponents, NaN, Inf, denormals), patterns with nearby exponents, and a grid of special values.
__device__ __forceinline__ unsigned long long sm64(unsigned long long& s) {
  unsigned long long z = (s += 0x9E3779B97F4A7C15ULL);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}
__global__ void div_selftest_kernel(unsigned long long seed, int per_thread, unsigned long long* mismatches) {
  const unsigned long long specials[12] = {0x0ULL, 0x8000000000000000ULL, 0x1ULL, 0x000FFFFFFFFFFFFFULL,
                                           0x0010000000000000ULL, 0x3FF0000000000000ULL, 0x7FEFFFFFFFFFFFFFULL,
                                           0x7FF0000000000000ULL, 0xFFF0000000000000ULL, 0x7FF8000000000000ULL,
                                           0x4059000000000000ULL, 0x3FEFFFFFFFFFFFFFULL};
  unsigned long long s = seed + 0x632BE59BD9B4E019ULL * (blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x + 1);
  unsigned long long bad = 0;
  for (int it = 0; it < per_thread; it++) {
    unsigned long long ua = sm64(s), ub = sm64(s);
    const int kind = it & 3;
    if (kind == 1) {  // exponents within +-64 of each other, mid range
      ua = (ua & 0x800FFFFFFFFFFFFFULL) | ((unsigned long long)(960 + (sm64(s) & 127)) << 52);
      ub = (ub & 0x800FFFFFFFFFFFFFULL) | ((unsigned long long)(960 + (sm64(s) & 127)) << 52);
    } else if (kind == 2) {  // one operand special
      if (it & 4) ua = specials[sm64(s) % 12]; else ub = specials[sm64(s) % 12];
    } else if (kind == 3) {  // mantissas with long runs of ones / zeros (hard cases for reciprocal refinement)
      ub = (ub & 0xFFF0000000000000ULL) | (0x000FFFFFFFFFFFFFULL >> (sm64(s) % 52)) ^ (sm64(s) & 3);
    }
    const double a = __longlong_as_double((long long)ua), b = __longlong_as_double((long long)ub);
    const Recip R = recip_prepare(b);
    const double got = (it & 8) ? div_by<true>(a, R) : div_by<false>(a, R), want = a / b;
    const unsigned long long ug = (unsigned long long)__double_as_longlong(got), uw = (unsigned long long)__double_as_longlong(want);
    const bool both_nan = (got != got) && (want != want);
    if (ug != uw && !both_nan) bad++;
    // the branch-free variants: whenever their deferred check passes, the value must be the IEEE one
    bool flag = false;
    const double gnb = (it & 8) ? div_nb<true>(a, R, flag) : div_nb<false>(a, R, flag);
    if (!flag && (unsigned long long)__double_as_longlong(gnb) != uw && !((gnb != gnb) && (want != want))) bad++;
    flag = false;
    const double snb = sqrt_nb(a, flag), swant = sqrt(a);
    if (!flag && (unsigned long long)__double_as_longlong(snb) != (unsigned long long)__double_as_longlong(swant)) bad++;
    if (kind == 1 && a > 0.0 && flag) bad++;  // mid-range positive arguments must take the fast path
  }
  if (bad) atomicAdd(mismatches, bad);
}
}  // namespace odeb200

extern "C" int64_t ode_b200_selftest_div(int device, int64_t n_samples, uint64_t seed) {
  using namespace odeb200;
  int cnt = 0;
  if (cudaGetDeviceCount(&cnt) != cudaSuccess || device < 0 || device >= cnt) {
    cudaGetLastError();
    return set_err(ODE_B200_E_NO_DEVICE, "no CUDA device %d", device);
  }
  ODEB200_CUDA(cudaSetDevice(device));
  unsigned long long* d = nullptr;
  ODEB200_CUDA(cudaMalloc(&d, 8));
  ODEB200_CUDA(cudaMemset(d, 0, 8));
  const int threads = 256, blocks = 148 * 8;
  long long per = (n_samples + (long long)threads * blocks - 1) / ((long long)threads * blocks);
  if (per < 1) per = 1;
  div_selftest_kernel<<<blocks, threads>>>(seed, (int)per, d);
  note_launch();
  ODEB200_CUDA(cudaGetLastError());
  unsigned long long bad = 0;
  ODEB200_CUDA(cudaMemcpy(&bad, d, 8, cudaMemcpyDeviceToHost));
  cudaFree(d);
  return (int64_t)bad;
}

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

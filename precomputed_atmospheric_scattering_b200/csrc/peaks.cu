// Roofline denominators of this path measured on the device itself: the LUT precompute is bound by
// the FP32 FMA pipe and, second, by the SFU (MUFU) pipe, for which the
// driver-written MEASURED_PEAKS.json has no figure. Two saturating microbenchmarks, timed with CUDA
// events: 16 independent FMA chains per thread (2 flops each), and 16 independent MUFU.RSQ chains.
#include <cuda_runtime.h>

#include <algorithm>
#include <string>

#include "../../include/pas_b200.h"

namespace {

constexpr int kChains = 16;
constexpr int kIters = 4096;

__global__ void __launch_bounds__(256) fma_peak_kernel(float* out, float a, float b) {
  float v[kChains];
#pragma unroll
  for (int i = 0; i < kChains; ++i) v[i] = (float)(threadIdx.x + i);
  for (int it = 0; it < kIters; ++it) {
#pragma unroll
    for (int i = 0; i < kChains; ++i) v[i] = fmaf(v[i], a, b);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kChains; ++i) s += v[i];
  if (s == 123.456f) out[0] = s;  // never true; keeps the chains alive
}

__global__ void __launch_bounds__(256) mufu_peak_kernel(float* out) {
  float v[kChains];
#pragma unroll
  for (int i = 0; i < kChains; ++i) v[i] = 1.5f + (float)(threadIdx.x + i);
  for (int it = 0; it < kIters; ++it) {
#pragma unroll
    for (int i = 0; i < kChains; ++i) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kChains; ++i) s += v[i];
  if (s == 123.456f) out[0] = s;
}

}  // namespace

extern "C" pas_status pas_measure_device_peaks(int device, double* fp32_tflops, double* mufu_gops,
                                               int* sm_count) {
  if (fp32_tflops == nullptr || mufu_gops == nullptr) return PAS_ERR_INVALID_ARGUMENT;
  if (cudaSetDevice(device) != cudaSuccess) return PAS_ERR_CUDA;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return PAS_ERR_CUDA;
  if (sm_count) *sm_count = prop.multiProcessorCount;
  float* out = nullptr;
  if (cudaMalloc(&out, 4) != cudaSuccess) return PAS_ERR_CUDA;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int blocks = prop.multiProcessorCount * 8, threads = 256;
  double best_fma = 0.0, best_mufu = 0.0;
  for (int rep = 0; rep < 6; ++rep) {
    float ms = 0.f;
    cudaEventRecord(e0);
    fma_peak_kernel<<<blocks, threads>>>(out, 0.999f, 0.001f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    const double flops = 2.0 * kChains * (double)kIters * blocks * threads;
    if (rep > 0 && ms > 0.f) best_fma = std::max(best_fma, flops / (ms * 1e-3) / 1e12);
    cudaEventRecord(e0);
    mufu_peak_kernel<<<blocks, threads>>>(out);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    const double ops = (double)kChains * kIters * blocks * threads;
    if (rep > 0 && ms > 0.f) best_mufu = std::max(best_mufu, ops / (ms * 1e-3) / 1e9);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  if (cudaGetLastError() != cudaSuccess) return PAS_ERR_CUDA;
  *fp32_tflops = best_fma;
  *mufu_gops = best_mufu;
  return PAS_OK;
}

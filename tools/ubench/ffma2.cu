// Microbenchmark: issue rate of packed fp32 FFMA2 vs scalar FFMA on sm_100a, alone and mixed with
// FADD.SAT (the instruction mix of the density kernel's azimuth sweep). Prints warp-instructions
// per clock per SM (scalar FFMA peak = 4). Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
// --fmad=false -o tools/ubench/ffma2 tools/ubench/ffma2.cu
#include <cuda_runtime.h>
#include <cstdio>

constexpr int kIters = 2048;

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, float a, float b) {
  float2 v[8];
  float w[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { v[i] = make_float2(threadIdx.x + i, threadIdx.x - i); w[i] = 0.01f * (threadIdx.x + i); }
  const float2 a2 = make_float2(a, a * 1.0001f), b2 = make_float2(b, b * 0.999f);
  for (int it = 0; it < kIters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) {  // 2 scalar FFMA
        v[i].x = fmaf(v[i].x, a, b);
        v[i].y = fmaf(v[i].y, a, b);
      } else if (MODE == 1) {  // 1 FFMA2
        v[i] = __ffma2_rn(v[i], a2, b2);
      } else if (MODE == 2) {  // 1 FFMA2 + 1 FADD.SAT
        v[i] = __ffma2_rn(v[i], a2, b2);
        w[i] = __saturatef(w[i] + a);
      } else if (MODE == 3) {  // 2 FFMA + 1 FADD.SAT
        v[i].x = fmaf(v[i].x, a, b);
        v[i].y = fmaf(v[i].y, a, b);
        w[i] = __saturatef(w[i] + a);
      } else if (MODE == 4) {  // 1 FFMA2 + 3 FADD.SAT (ramp pair of the sweep: per 2 FMAs 3 adds)
        v[i] = __ffma2_rn(v[i], a2, b2);
        w[i] = __saturatef(w[i] + a);
        w[(i + 1) & 7] = __saturatef(w[(i + 1) & 7] + b);
        w[(i + 2) & 7] = __saturatef(w[(i + 2) & 7] - a);
      } else if (MODE == 5) {  // 2 FFMA + 3 FADD.SAT
        v[i].x = fmaf(v[i].x, a, b);
        v[i].y = fmaf(v[i].y, a, b);
        w[i] = __saturatef(w[i] + a);
        w[(i + 1) & 7] = __saturatef(w[(i + 1) & 7] + b);
        w[(i + 2) & 7] = __saturatef(w[(i + 2) & 7] - a);
      } else if (MODE == 6) {  // FFMA2 + FMNMX pair (ALU pipe) : does the ALU pipe co-issue?
        v[i] = __ffma2_rn(v[i], a2, b2);
        w[i] = fminf(fmaxf(w[i], a), b);
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += v[i].x + v[i].y + w[i];
  if (s == 123.456f) out[0] = s;
}

template <int MODE>
void run(const char* name, double inst_per_iter, double fma_flops_per_iter) {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  float* out;
  cudaMalloc(&out, 4);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int blocks = prop.multiProcessorCount * 8, threads = 256;
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(out, 0.999f, 0.001f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
  }
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double warps = (double)blocks * threads / 32;
  const double inst = warps * kIters * 8 * inst_per_iter;
  const double clocks = best * 1e-3 * khz * 1e3;
  printf("%-34s %8.3f ms  %6.3f warp-inst/clk/SM (at %d MHz nominal)  %7.2f TFLOP/s fp32\n", name, best,
         inst / clocks / prop.multiProcessorCount, khz / 1000,
         warps * 32 * kIters * 8 * fma_flops_per_iter / (best * 1e-3) / 1e12);
  cudaFree(out);
}

int main() {
  run<0>("2 FFMA", 2, 4);
  run<1>("1 FFMA2", 1, 4);
  run<2>("1 FFMA2 + 1 FADD.SAT", 2, 4);
  run<3>("2 FFMA + 1 FADD.SAT", 3, 4);
  run<4>("1 FFMA2 + 3 FADD.SAT", 4, 4);
  run<5>("2 FFMA + 3 FADD.SAT", 5, 4);
  run<6>("1 FFMA2 + 2 FMNMX", 3, 4);
  return 0;
}

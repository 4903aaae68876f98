// Batched render-time lookups: kernels over the device functions of
// kernel_render.cuh, one thread per query, behind pas_model_get_sky_radiance /
// pas_model_get_sky_radiance_to_point / pas_model_get_sun_and_sky_irradiance. The tables (T 256 KiB,
// E 16 KiB, S 8-16 MiB) live in L2. Vectors are [n][3] doubles, outputs [n][3] floats.
#include "kernel_render.cuh"

namespace pas {
namespace {

__device__ __forceinline__ void store3(float* out, size_t q, V3 v) {
  out[3 * q + 0] = (float)v.x;
  out[3 * q + 1] = (float)v.y;
  out[3 * q + 2] = (float)v.z;
}

__global__ void __launch_bounds__(128)
sky_radiance_kernel(const __grid_constant__ RenderContext k, size_t n, int to_point,
                    const double* __restrict__ camera, const double* __restrict__ target,
                    const double* __restrict__ shadow_length, const double* __restrict__ sun_direction,
                    float* __restrict__ radiance, float* __restrict__ transmittance) {
  const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n) return;
  V3 tr;
  const double shadow = shadow_length != nullptr ? shadow_length[q] : 0.0;
  const V3 cam = load3(camera + 3 * q), tgt = load3(target + 3 * q), sun = load3(sun_direction + 3 * q);
  const V3 L = to_point ? sky_radiance_to_point(k, cam, tgt, shadow, sun, &tr) : sky_radiance(k, cam, tgt, shadow, sun, &tr);
  store3(radiance, q, L);
  if (transmittance != nullptr) store3(transmittance, q, tr);
}

__global__ void __launch_bounds__(128)
sun_and_sky_irradiance_kernel(const __grid_constant__ RenderContext k, size_t n, const double* __restrict__ point,
                              const double* __restrict__ normal, const double* __restrict__ sun_direction,
                              float* __restrict__ sun_irradiance, float* __restrict__ sky_irradiance) {
  const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n) return;
  V3 sky;
  const V3 sun = sun_and_sky_irradiance(k, load3(point + 3 * q), load3(normal + 3 * q), load3(sun_direction + 3 * q), &sky);
  store3(sun_irradiance, q, sun);
  store3(sky_irradiance, q, sky);
}

}  // namespace

cudaError_t launch_sky_radiance(const PasGeometry& g, const RenderTables& t, const RenderConstants& c,
                                size_t n, bool to_point, const double* camera, const double* target,
                                const double* shadow_length, const double* sun_direction, float* radiance,
                                float* transmittance, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  sky_radiance_kernel<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(
      RenderContext{g, t, c}, n, to_point ? 1 : 0, camera, target, shadow_length, sun_direction, radiance, transmittance);
  return cudaGetLastError();
}

cudaError_t launch_sun_and_sky_irradiance(const PasGeometry& g, const RenderTables& t,
                                          const RenderConstants& c, size_t n, const double* point,
                                          const double* normal, const double* sun_direction,
                                          float* sun_irradiance, float* sky_irradiance, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  sun_and_sky_irradiance_kernel<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(
      RenderContext{g, t, c}, n, point, normal, sun_direction, sun_irradiance, sky_irradiance);
  return cudaGetLastError();
}

}  // namespace pas

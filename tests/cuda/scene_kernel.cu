// TEST INFRASTRUCTURE -- not part of libpas_b200.so. The scene of the reference's integration test
// (atmosphere/reference/model_test.glsl:66-348: a sphere standing on a spherical planet, lit by the sun
// and the sky, with its shadow volume / light shafts; view rays and tone map of
// atmosphere/reference/model_test.cc:688-736) as a CUDA kernel over the product's device-side lookups
// (csrc/kernel_render.cuh, context from pas_model_render_context). The tests compare its images with the
// oracle's renderer and with golden images of the unmodified reference, with the reference's own PSNR
// thresholds: that pins sky_radiance / sky_radiance_to_point / sun_and_sky_irradiance in the
// configurations the reference tests them in. Built by tests/cuda/Makefile into
// tests/cuda/libpas_test_scene.so, loaded by tests/scene_render.py.
#include <cuda_runtime.h>

#include <cstdint>

#include "../../precomputed_atmospheric_scattering_b200/csrc/kernel_render.cuh"

namespace {

using namespace pas;

struct SceneView {  // uniforms of reference/model_test.cc:127-134 + image size (tests/scene.py: View)
  double camera[3], earth_center[3], sun_direction[3], sun_size[2];
  double sphere_center[3], sphere_radius;
  double model_from_clip[9];
  double ground_albedo[3], sphere_albedo[3];
  double exposure;
  int use_luminance, width, height;
};

// The umbra / penumbra cone of the sphere cut by a view ray: [d_in, d_out] (model_test.glsl:151-191).
struct Interval {
  double lo, hi;
};
__device__ Interval shadow_cone_interval(const SceneView& s, V3 dir) {
  const V3 o = load3(s.camera) - load3(s.sphere_center), sun = load3(s.sun_direction);
  const double k = s.sun_size[0], R = s.sphere_radius, widen = 1.0 + k * k;
  const double o_sun = dot(o, sun), d_sun = dot(dir, sun);
  // quadratic a d^2 + 2 b d + c = 0 of the cone surface
  const double a = 1.0 - widen * d_sun * d_sun;
  const double b = dot(o, dir) - widen * o_sun * d_sun - k * R * d_sun;
  const double c = dot(o, o) - widen * o_sun * o_sun - 2.0 * k * R * o_sun - R * R;
  const double disc = b * b - a * c;
  if (!(disc > 0.0)) return Interval{0.0, 0.0};
  const double root = sqrt(disc);
  Interval iv{fmax(0.0, (-b - root) / a), (-b + root) / a};
  // keep the half of the cone behind the sphere, between its base plane and its apex
  const double base = -o_sun / d_sun, apex = -(o_sun + R / k) / d_sun;
  if (d_sun > 0.0) {
    iv.lo = fmax(iv.lo, apex);
    iv.hi = a > 0.0 ? fmin(iv.hi, base) : base;
  } else {
    iv.lo = a > 0.0 ? fmax(iv.lo, base) : base;
    iv.hi = fmin(iv.hi, apex);
  }
  return iv;
}

// First intersection of the ray (origin o relative to the sphere centre, unit direction) with a sphere;
// also the grazing distance used for anti-aliasing (model_test.glsl:232-247). Negative: no hit.
struct SphereHit {
  double distance, graze;
};
__device__ SphereHit hit_sphere(V3 o, V3 dir, double radius) {
  const double od = dot(o, dir), perp2 = dot(o, o) - od * od;
  return SphereHit{-od - sqrt(radius * radius - perp2), (radius - sqrt(perp2)) / -od};
}

// Light reflected by a Lambertian surface point towards the camera, seen through the atmosphere
// (model_test.glsl:253-275, 287-313).
__device__ V3 lit_surface(const RenderContext& k, const SceneView& s, V3 point, V3 normal, V3 albedo,
                          double sun_factor, double sky_factor, double in_shadow) {
  const V3 centre = load3(s.earth_center), sun = load3(s.sun_direction);
  V3 sky;
  const V3 direct = sun_and_sky_irradiance(k, point - centre, normal, sun, &sky);
  V3 through;
  const V3 haze = sky_radiance_to_point(k, load3(s.camera) - centre, point - centre, in_shadow, sun, &through);
  return albedo * (1.0 / kPi) * (direct * sun_factor + sky * sky_factor) * through + haze;
}

// GetViewRayRadiance (model_test.glsl:218-348) for one pixel.
__device__ V3 pixel_radiance(const RenderContext& k, const SceneView& s, V3 ray, V3 ray_step) {
  const double len = norm(ray);
  const V3 dir = ray * (1.0 / len);
  const double pixel_angle = norm(ray_step) / len;
  const Interval cone = shadow_cone_interval(s, dir);
  const V3 camera = load3(s.camera), centre = load3(s.earth_center), ball = load3(s.sphere_center), sun = load3(s.sun_direction);

  // background: sky (+ sun disc)
  V3 to_space;
  V3 colour = sky_radiance(k, camera - centre, dir, fmax(0.0, cone.hi - cone.lo), sun, &to_space);
  if (dot(dir, sun) > s.sun_size[1]) colour = colour + to_space * solar_radiance(k);

  // the planet
  const SphereHit planet = hit_sphere(camera - centre, dir, fabs(centre.z));
  if (planet.distance > 0.0) {
    const V3 point = camera + dir * planet.distance;
    V3 normal = point - centre;
    normal = normal * (1.0 / norm(normal));
    // the sphere hides part of the sun and of the sky from the ground (model_test.glsl:100-139)
    const V3 q = point - ball;
    const SphereHit towards_sun = hit_sphere(q, sun, s.sphere_radius);
    const double sun_seen = towards_sun.distance > 0.0
                                ? render_detail::ramp01(1.0, 0.0, towards_sun.graze / s.sun_size[0]) : 1.0;
    const double qq = dot(q, q);
    const double sky_seen = 1.0 + q.z / sqrt(qq) * s.sphere_radius * s.sphere_radius / qq;
    const double in_shadow = fmax(0.0, fmin(cone.hi, planet.distance) - cone.lo);
    colour = lit_surface(k, s, point, normal, load3(s.ground_albedo), sun_seen, sky_seen, in_shadow);
  }

  // the sphere, anti-aliased at its silhouette
  const SphereHit sphere = hit_sphere(camera - ball, dir, s.sphere_radius);
  if (sphere.distance > 0.0) {
    const double coverage = fmin(sphere.graze / pixel_angle, 1.0);
    const V3 point = camera + dir * sphere.distance;
    V3 normal = point - ball;
    normal = normal * (1.0 / norm(normal));
    const double in_shadow = fmax(0.0, fmin(cone.hi, sphere.distance) - cone.lo);
    const V3 lit = lit_surface(k, s, point, normal, load3(s.sphere_albedo), 1.0, 1.0, in_shadow);
    colour = colour * (1.0 - coverage) + lit * coverage;
  }
  return colour;
}

__device__ __forceinline__ unsigned tone(double v, double exposure) {
  // model_test.cc:726-731: pow(1 - exp(-v * exposure), 1 / 2.2), truncated to 8 bits
  return (unsigned)(pow(1.0 - exp(-v * exposure), 1.0 / 2.2) * 255.0);
}

__global__ void __launch_bounds__(128)
scene_kernel(const __grid_constant__ RenderContext k, const __grid_constant__ SceneView s, float* __restrict__ rgb,
             unsigned* __restrict__ argb) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y * blockDim.y + threadIdx.y;
  if (i >= s.width || j >= s.height) return;
  // view ray of pixel (i, j), j = 0 at the top, and its change to the next pixel (model_test.cc:690-711)
  const double x = 2.0 * (i + 0.5) / s.width - 1.0, y = 1.0 - 2.0 * (j + 0.5) / s.height;
  const double dx = 2.0 / s.width, dy = -2.0 / s.height;
  const double* M = s.model_from_clip;
  const V3 ray = v3(M[0] * x + M[1] * y + M[2], M[3] * x + M[4] * y + M[5], M[6] * x + M[7] * y + M[8]);
  const V3 step = v3(M[0] * dx + M[1] * dy, M[3] * dx + M[4] * dy, M[6] * dx + M[7] * dy);
  const V3 L = pixel_radiance(k, s, ray, step);
  const size_t p = (size_t)j * s.width + i;
  if (rgb != nullptr) {
    rgb[3 * p + 0] = (float)L.x;
    rgb[3 * p + 1] = (float)L.y;
    rgb[3 * p + 2] = (float)L.z;
  }
  if (argb != nullptr) {
    argb[p] = (255u << 24) | (tone(L.x, s.exposure) << 16) | (tone(L.y, s.exposure) << 8) | tone(L.z, s.exposure);
  }
}

}  // namespace

// context: the bytes pas_model_render_context() filled; view: SceneView; rgb / argb: HOST buffers (either
// may be NULL). Returns 0 or a CUDA error code; *kernel_ms = device time of the kernel.
extern "C" int pas_test_render_scene(const void* context, size_t context_bytes, const void* view, size_t view_bytes,
                                     float* rgb, uint32_t* argb, float* kernel_ms) {
  if (context_bytes != sizeof(RenderContext) || view_bytes != sizeof(SceneView)) return -1;
  const RenderContext k = *static_cast<const RenderContext*>(context);
  const SceneView s = *static_cast<const SceneView*>(view);
  const size_t pixels = (size_t)s.width * s.height;
  float* d_rgb = nullptr;
  unsigned* d_argb = nullptr;
  cudaError_t e = cudaSuccess;
  if (rgb != nullptr) e = cudaMalloc(&d_rgb, pixels * 3 * sizeof(float));
  if (e == cudaSuccess && argb != nullptr) e = cudaMalloc(&d_argb, pixels * sizeof(unsigned));
  cudaEvent_t t0 = nullptr, t1 = nullptr;
  if (e == cudaSuccess) e = cudaEventCreate(&t0);
  if (e == cudaSuccess) e = cudaEventCreate(&t1);
  if (e == cudaSuccess) {
    const dim3 block(32, 4), grid((s.width + 31) / 32, (s.height + 3) / 4);
    cudaEventRecord(t0);
    scene_kernel<<<grid, block>>>(k, s, d_rgb, d_argb);
    cudaEventRecord(t1);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess && rgb != nullptr) e = cudaMemcpy(rgb, d_rgb, pixels * 3 * sizeof(float), cudaMemcpyDeviceToHost);
  if (e == cudaSuccess && argb != nullptr) e = cudaMemcpy(argb, d_argb, pixels * sizeof(unsigned), cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) e = cudaEventSynchronize(t1);
  if (e == cudaSuccess && kernel_ms != nullptr) cudaEventElapsedTime(kernel_ms, t0, t1);
  if (t0) cudaEventDestroy(t0);
  if (t1) cudaEventDestroy(t1);
  cudaFree(d_rgb);
  cudaFree(d_argb);
  return (int)e;
}

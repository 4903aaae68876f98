// Device-side lookups in the precomputed tables: the CUDA counterpart of the GLSL the reference hands
// to its users (atmosphere::Model::shader(), atmosphere/model.cc:221-281). A CUDA renderer includes this
// header, obtains a pas::RenderContext for a model with pas_model_render_context() (include/pas_b200.h)
// and calls
//   pas::sky_radiance            GetSkyRadiance / GetSkyLuminance            functions.glsl:1705-1769
//   pas::sky_radiance_to_point   GetSkyRadianceToPoint / ...LuminanceToPoint functions.glsl:1787-1863
//   pas::sun_and_sky_irradiance  GetSunAndSkyIrradiance / ...Illuminance     functions.glsl:1878-1896
//   pas::solar_radiance          GetSolarRadiance / GetSolarLuminance        model.cc:228-231, 254-258
// from its own kernels; the library's batched entry points (pas_model_get_sky_radiance, ...) are
// kernels over the same functions (kernel_render.cu).
//
// Organisation (not the GLSL's): a view ray is reduced once to a `Ray` -- origin moved to the top of the
// atmosphere if it starts outside, (r, mu, mu_s, nu), ground flag, and the transmittance to the boundary
// at the origin, which every transmittance along the ray divides by or into -- and every in-scattered
// radiance is one `inscatter()`: a single 4-D footprint (2 nu slabs x 8 corners, weights multiplied out
// once) gathered from the scattering table and, if present, the single-Mie table together.
//
// Numerics: a view ray starts ~6360 km from the planet centre and the quantities that select table
// texels are differences of such lengths: the geometry runs in double (B200 has the fp64 rate for it,
// a query needs ~2k flops), the table texels are fp32 / fp16. Fetches are software bi/tri-linear with
// the CPU reference's index / weight / clamp rule (dimensional_types binary_function.h:103-118,
// ternary_function.h:100-125), not 8-bit-fraction hardware filtering.
#ifndef PAS_B200_CSRC_KERNEL_RENDER_CUH_
#define PAS_B200_CSRC_KERNEL_RENDER_CUH_

#include "pas_kernels.h"
#include "pas_physics.cuh"

namespace pas {

// Everything the lookups need, by value (a kernel argument): filled by pas_model_render_context().
struct RenderContext {
  PasGeometry g;
  RenderTables t;
  RenderConstants c;
};

struct V3 {
  double x, y, z;
};
__device__ __forceinline__ V3 v3(double x, double y, double z) { return V3{x, y, z}; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 operator*(V3 a, V3 b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }
__device__ __forceinline__ V3 operator*(V3 a, double s) { return v3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ V3 operator/(V3 a, V3 b) { return v3(a.x / b.x, a.y / b.y, a.z / b.z); }
__device__ __forceinline__ double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ double norm(V3 a) { return sqrt(dot(a, a)); }
__device__ __forceinline__ V3 load3(const double* p) { return v3(p[0], p[1], p[2]); }
__device__ __forceinline__ V3 splat(double s) { return v3(s, s, s); }
__device__ __forceinline__ V3 at_most_one(V3 a) { return v3(fmin(a.x, 1.0), fmin(a.y, 1.0), fmin(a.z, 1.0)); }

namespace render_detail {

struct Tap1 {  // one table axis: the two texels around a texel-space coordinate and the weight of the second
  int lo, hi;
  double w;
};
__device__ __forceinline__ Tap1 tap1(double x, int n) {
  const double f = floor(x);
  const int i = (int)f;
  return Tap1{min(max(i, 0), n - 1), min(max(i + 1, 0), n - 1), x - f};
}

// bilinear fetch in an RGBA32F 2-D table
__device__ inline V3 fetch_2d(const float4* tab, int width, Tap1 x, Tap1 y) {
  double acc[3] = {0.0, 0.0, 0.0};
#pragma unroll
  for (int corner = 0; corner < 4; ++corner) {
    const float4 v = __ldg(tab + ((corner & 1) ? x.hi : x.lo) + (size_t)width * ((corner & 2) ? y.hi : y.lo));
    const double w = ((corner & 1) ? x.w : 1.0 - x.w) * ((corner & 2) ? y.w : 1.0 - y.w);
    acc[0] += w * v.x;
    acc[1] += w * v.y;
    acc[2] += w * v.z;
  }
  return v3(acc[0], acc[1], acc[2]);
}

// one RGBA texel of a 3-D table, fp32 or fp16 storage
__device__ __forceinline__ float4 texel_3d(const void* tab, int half, size_t i) {
  if (half) {
    const uint2 raw = __ldg(reinterpret_cast<const uint2*>(tab) + i);
    const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&raw.x));
    const float2 hi = __half22float2(*reinterpret_cast<const __half2*>(&raw.y));
    return make_float4(lo.x, lo.y, hi.x, hi.y);
  }
  return __ldg(reinterpret_cast<const float4*>(tab) + i);
}

__device__ __forceinline__ double ramp01(double lo, double hi, double x) {  // smoothstep
  const double t = d_clamp((x - lo) / (hi - lo), 0.0, 1.0);
  return t * t * (3.0 - 2.0 * t);
}

}  // namespace render_detail

// Transmittance from radius r to the top boundary along a direction of zenith cosine mu
// (GetTransmittanceToTopAtmosphereBoundary, functions.glsl:473-480).
__device__ inline V3 transmittance_to_top(const RenderContext& k, double r, double mu) {
  double x, y;
  transmittance_xy(k.g, r, mu, &x, &y);
  return render_detail::fetch_2d(k.t.transmittance, k.g.sz.t_w, render_detail::tap1(x, k.g.sz.t_w),
                                 render_detail::tap1(y, k.g.sz.t_h));
}

// A view ray inside the atmosphere.
struct Ray {
  double r, mu, mu_s, nu;  // at the origin
  bool hit;                // reaches the ground
  bool outside;            // never enters the atmosphere
  double entered;          // distance from the camera to the origin (> 0 for a camera in space)
  V3 t_origin;             // transmittance to the boundary from the origin, along the ray (hit: against it)
};

// Reduces (camera, unit view direction, sun direction) -- camera relative to the planet centre -- to a Ray;
// a camera in space is moved to where the ray enters the atmosphere (functions.glsl:1713-1727, 1795-1807).
__device__ inline Ray make_ray(const RenderContext& k, V3 camera, V3 view, V3 sun) {
  Ray ray;
  double r = norm(camera), r_mu = dot(camera, view);
  const double to_top = -r_mu - sqrt(r_mu * r_mu - r * r + k.g.top * k.g.top);
  ray.outside = false;
  ray.entered = 0.0;
  if (to_top > 0.0) {
    camera = camera + view * to_top;
    r = k.g.top;
    r_mu += to_top;
    ray.entered = to_top;
  } else if (r > k.g.top) {
    ray.outside = true;
  }
  ray.r = r;
  ray.mu = r_mu / r;
  ray.mu_s = dot(camera, sun) / r;
  ray.nu = dot(view, sun);
  ray.hit = hits_ground(k.g, r, ray.mu);
  ray.t_origin = transmittance_to_top(k, r, ray.hit ? -ray.mu : ray.mu);
  return ray;
}

// The point at distance d along the ray: radius and the two cosines there.
struct RayPoint {
  double r, mu, mu_s;
};
__device__ __forceinline__ RayPoint along(const RenderContext& k, const Ray& ray, double d) {
  RayPoint p;
  p.r = d_clamp(sqrt(d * d + 2.0 * ray.r * ray.mu * d + ray.r * ray.r), k.g.bottom, k.g.top);
  p.mu = (ray.r * ray.mu + d) / p.r;
  p.mu_s = (ray.r * ray.mu_s + d * ray.nu) / p.r;
  return p;
}

// Transmittance between the origin and the point at distance d (GetTransmittance, functions.glsl:493-519):
// a ratio of two boundary transmittances, the origin's being the ray's.
__device__ inline V3 transmittance_along(const RenderContext& k, const Ray& ray, double d) {
  const RayPoint p = along(k, ray, d);
  const double mu_d = d_clamp(p.mu, -1.0, 1.0);
  return ray.hit ? at_most_one(transmittance_to_top(k, p.r, -mu_d) / ray.t_origin)
                 : at_most_one(ray.t_origin / transmittance_to_top(k, p.r, mu_d));
}

// In-scattered radiance table lookup at (r, mu, mu_s, nu): Rayleigh + multiple scattering in the return
// value, single Mie scattering in *mie (GetCombinedScattering, functions.glsl:1658-1690, with the
// extrapolation of functions.glsl:1634-1646 when the Mie term is packed in the alpha channel). One
// footprint of 2 nu slabs x (2 x 2 x 2) corners for both tables.
__device__ inline V3 inscatter(const RenderContext& k, double r, double mu, double mu_s, double nu, bool hit, V3* mie) {
  using namespace render_detail;
  const PasSizes& z = k.g.sz;
  const int width = z.nu_n * z.mu_s_n;
  // forward mapping (functions.glsl:773-831) in texel space
  const double rho = sqrt(d_pos(r * r - k.g.bottom * k.g.bottom));
  const Tap1 layer = tap1(rho / k.g.H * (z.r_n - 1), z.r_n);
  const Tap1 row = tap1(scattering_y_from_mu(k.g, r, rho, mu, hit), z.mu_n);
  const double col = scattering_x_from_mu_s(k.g, mu_s);
  const double slab_x = (nu + 1.0) * 0.5 * (z.nu_n - 1), slab = floor(slab_x);
  const Tap1 cols[2] = {tap1(slab * z.mu_s_n + col, width), tap1((slab + 1.0) * z.mu_s_n + col, width)};
  const double slab_w[2] = {1.0 - (slab_x - slab), slab_x - slab};
  const bool packed = k.t.single_mie == nullptr;
  double s[4] = {0.0, 0.0, 0.0, 0.0}, m[3] = {0.0, 0.0, 0.0};
#pragma unroll
  for (int corner = 0; corner < 16; ++corner) {
    const Tap1& cx = cols[corner >> 3];
    const double w = slab_w[corner >> 3] * ((corner & 1) ? cx.w : 1.0 - cx.w) *
                     ((corner & 2) ? row.w : 1.0 - row.w) * ((corner & 4) ? layer.w : 1.0 - layer.w);
    const size_t texel = ((corner & 1) ? cx.hi : cx.lo) +
                         (size_t)width * (((corner & 2) ? row.hi : row.lo) + (size_t)z.mu_n * ((corner & 4) ? layer.hi : layer.lo));
    const float4 v = texel_3d(k.t.scattering, k.t.half_precision, texel);
    s[0] += w * v.x; s[1] += w * v.y; s[2] += w * v.z; s[3] += w * v.w;
    if (!packed) {
      const float4 u = texel_3d(k.t.single_mie, k.t.half_precision, texel);
      m[0] += w * u.x; m[1] += w * u.y; m[2] += w * u.z;
    }
  }
  *mie = packed ? v3(s[3], 0.0, 0.0) : v3(m[0], m[1], m[2]);
  return v3(s[0], s[1], s[2]);
}

// Single Mie scattering from its red channel and the Rayleigh + multiple term
// (GetExtrapolatedSingleMieScattering, functions.glsl:1634-1646).
__device__ inline V3 unpack_mie(const RenderContext& k, V3 scattering, double mie_red) {
  if (scattering.x <= 0.0) return splat(0.0);
  const double f = mie_red / scattering.x * (k.c.rayleigh[0] / k.c.mie_sca[0]);
  return v3(scattering.x * f * (k.c.mie_sca[0] / k.c.rayleigh[0]), scattering.y * f * (k.c.mie_sca[1] / k.c.rayleigh[1]),
            scattering.z * f * (k.c.mie_sca[2] / k.c.rayleigh[2]));
}

// scattering x Rayleigh phase + single Mie x Cornette-Shanks phase, in the units of the context
// (radiance, or luminance with SKY_SPECTRAL_RADIANCE_TO_LUMINANCE)
__device__ inline V3 apply_phases(const RenderContext& k, double nu, V3 scattering, V3 mie) {
  const double g = k.g.mie_g;
  const double p_mie = mie_phase_k(g) * (1.0 + nu * nu) / pow(1.0 + g * g - 2.0 * g * nu, 1.5);
  return (scattering * rayleigh_phase(nu) + mie * p_mie) * load3(k.c.sky_k);
}

// GetSkyRadiance: radiance arriving at `camera` (relative to the planet centre) from direction `view`,
// with the first `shadow_length` of the ray in shadow; *transmittance = to the top of the atmosphere.
__device__ inline V3 sky_radiance(const RenderContext& k, V3 camera, V3 view, double shadow_length, V3 sun,
                                  V3* transmittance) {
  const Ray ray = make_ray(k, camera, view, sun);
  if (ray.outside) {
    *transmittance = splat(1.0);
    return splat(0.0);
  }
  *transmittance = ray.hit ? splat(0.0) : ray.t_origin;
  V3 scattering, mie;
  if (shadow_length == 0.0) {
    scattering = inscatter(k, ray.r, ray.mu, ray.mu_s, ray.nu, ray.hit, &mie);
    if (k.t.single_mie == nullptr) mie = unpack_mie(k, scattering, mie.x);
  } else {
    // light shafts: what the table holds beyond the shadowed stretch, attenuated on the way back
    const RayPoint p = along(k, ray, shadow_length);
    scattering = inscatter(k, p.r, p.mu, p.mu_s, ray.nu, ray.hit, &mie);
    if (k.t.single_mie == nullptr) mie = unpack_mie(k, scattering, mie.x);
    const V3 t = transmittance_along(k, ray, shadow_length);
    scattering = scattering * t;
    mie = mie * t;
  }
  return apply_phases(k, ray.nu, scattering, mie);
}

// GetSkyRadianceToPoint: radiance in-scattered between `camera` and `point`, the last `shadow_length`
// of the segment in shadow; *transmittance = between the two.
__device__ inline V3 sky_radiance_to_point(const RenderContext& k, V3 camera, V3 point, double shadow_length, V3 sun,
                                           V3* transmittance) {
  const V3 offset = point - camera;
  const double length = norm(offset);
  const Ray ray = make_ray(k, camera, offset * (1.0 / length), sun);
  // the segment is measured from where the ray enters the atmosphere (functions.glsl:1812)
  const double d = fabs(length - ray.entered);
  *transmittance = transmittance_along(k, ray, d);
  V3 mie_near, mie_far;
  const V3 near = inscatter(k, ray.r, ray.mu, ray.mu_s, ray.nu, ray.hit, &mie_near);
  const double d_lit = fmax(d - shadow_length, 0.0);
  const RayPoint p = along(k, ray, d_lit);
  const V3 far = inscatter(k, p.r, p.mu, p.mu_s, ray.nu, ray.hit, &mie_far);
  if (k.t.single_mie == nullptr) {
    // a packed Mie term counts only where there is Rayleigh light to scale (functions.glsl:1640-1642)
    if (near.x <= 0.0) mie_near.x = 0.0;
    if (far.x <= 0.0) mie_far.x = 0.0;
  }
  const V3 t_lit = shadow_length > 0.0 ? transmittance_along(k, ray, d_lit) : *transmittance;
  // segment = (origin .. infinity) - transmittance x (far point .. infinity); the packed Mie term is
  // unpacked after the subtraction (functions.glsl:1845-1854)
  const V3 scattering = near - t_lit * far;
  V3 mie;
  if (k.t.single_mie == nullptr) {
    mie = unpack_mie(k, scattering, mie_near.x - t_lit.x * mie_far.x);
  } else {
    mie = mie_near - t_lit * mie_far;
  }
  mie = mie * render_detail::ramp01(0.0, 0.01, ray.mu_s);  // hack of functions.glsl:1856-1858
  return apply_phases(k, ray.nu, scattering, mie);
}

// GetSunAndSkyIrradiance: direct sun irradiance on a surface of the given normal at `point` (return value)
// and sky irradiance (*sky), GetIrradiance (functions.glsl:1524-1533, 1595-1601) and
// GetTransmittanceToSun (functions.glsl:552-563) inside.
__device__ inline V3 sun_and_sky_irradiance(const RenderContext& k, V3 point, V3 normal, V3 sun, V3* sky) {
  using namespace render_detail;
  const double r = norm(point), mu_s = dot(point, sun) / r;
  const PasSizes& z = k.g.sz;
  const V3 ground = fetch_2d(k.t.irradiance, z.e_w, tap1((mu_s * 0.5 + 0.5) * (z.e_w - 1), z.e_w),
                             tap1((r - k.g.bottom) / (k.g.top - k.g.bottom) * (z.e_h - 1), z.e_h));
  *sky = ground * ((1.0 + dot(normal, point) / r) * 0.5) * load3(k.c.sky_k);
  // fraction of the sun disc above the horizon
  const double sin_h = k.g.bottom / r, cos_h = -sqrt(d_pos(1.0 - sin_h * sin_h)), a = k.g.sun_angular_radius;
  const V3 to_sun = transmittance_to_top(k, r, mu_s) * ramp01(-sin_h * a, sin_h * a, mu_s - cos_h);
  return load3(k.c.solar) * to_sun * fmax(dot(normal, sun), 0.0) * load3(k.c.sun_k);
}

// GetSolarRadiance / GetSolarLuminance (model.cc:228-231, 254-258).
__device__ inline V3 solar_radiance(const RenderContext& k) {
  const double a = k.g.sun_angular_radius;
  return load3(k.c.solar) * (1.0 / (kPi * a * a)) * load3(k.c.sun_k);
}

}  // namespace pas

#endif  // PAS_B200_CSRC_KERNEL_RENDER_CUH_

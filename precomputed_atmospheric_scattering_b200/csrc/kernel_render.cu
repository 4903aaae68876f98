// Render-time use of the precomputed tables: the three public
// lookups of the rendering API
//   GetSkyRadiance           atmosphere/functions.glsl:1705-1769
//   GetSkyRadianceToPoint    atmosphere/functions.glsl:1787-1863
//   GetSunAndSkyIrradiance   atmosphere/functions.glsl:1878-1896
// with GetCombinedScattering / GetExtrapolatedSingleMieScattering (functions.glsl:1634-1690), their
// luminance wrappers (atmosphere/model.cc:221-281), and the integration-test scene
// GetViewRayRadiance of atmosphere/reference/model_test.glsl:66-348 with the view rays and tone map
// of reference/model_test.cc:688-736.
//
// Numerics: a view ray starts ~6360 km from the planet centre and the quantities that select table
// texels are differences of such lengths (r - bottom, d - d_min, ...): in fp32 (what the GLSL
// renderer uses) they keep 3-4 digits near the ground. B200 has a full-rate fp64 pipe, and a pixel
// needs only ~2k flops, so the geometry runs in double and only the table texels are fp32 / fp16:
// the images then match the fp64 CPU model to ~1e-6 wherever the tables do. Table fetches are
// software bi/tri-linear with the CPU reference's index / weight / clamp rule
// (dimensional_types binary_function.h:103-118, ternary_function.h:100-125), not 8-bit-fraction
// hardware filtering. One thread per pixel (or per query); the tables (T 256 KiB, E 16 KiB,
// S 8-16 MiB) live in L2.
#include "pas_kernels.h"
#include "pas_physics.cuh"

namespace pas {
namespace {

struct V3 {
  double x, y, z;
};
__device__ __forceinline__ V3 v3(double x, double y, double z) { return V3{x, y, z}; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 operator*(V3 a, V3 b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }
__device__ __forceinline__ V3 operator*(V3 a, double s) { return v3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ V3 load3(const double* p) { return v3(p[0], p[1], p[2]); }
__device__ __forceinline__ V3 splat(double s) { return v3(s, s, s); }

__device__ __forceinline__ V3 texel_rgb(const float4* t, size_t i) {
  const float4 v = __ldg(t + i);
  return v3(v.x, v.y, v.z);
}
struct Rgba {
  double r, g, b, a;
};
__device__ __forceinline__ Rgba texel_rgba(const void* t, int half, size_t i) {
  if (half) {
    const uint2 raw = __ldg(reinterpret_cast<const uint2*>(t) + i);
    const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&raw.x));
    const float2 hi = __half22float2(*reinterpret_cast<const __half2*>(&raw.y));
    return Rgba{lo.x, lo.y, hi.x, hi.y};
  }
  const float4 v = __ldg(reinterpret_cast<const float4*>(t) + i);
  return Rgba{v.x, v.y, v.z, v.w};
}

struct TapD {
  int i0, i1;
  double w;
};
__device__ __forceinline__ TapD tap_d(double x, int n) {  // x in texel space (u * n - 0.5)
  TapD t;
  const double fl = floor(x);
  const int i = (int)fl;
  t.w = x - fl;
  t.i0 = min(max(i, 0), n - 1);
  t.i1 = min(max(i + 1, 0), n - 1);
  return t;
}

struct Ctx {
  const PasGeometry& g;
  const RenderTables& t;
  const RenderConstants& c;
};

__device__ V3 fetch2(const float4* tab, int w, TapD tx, TapD ty) {
  const V3 a = texel_rgb(tab, tx.i0 + (size_t)w * ty.i0), b = texel_rgb(tab, tx.i1 + (size_t)w * ty.i0);
  const V3 e = texel_rgb(tab, tx.i0 + (size_t)w * ty.i1), d = texel_rgb(tab, tx.i1 + (size_t)w * ty.i1);
  const double wx = tx.w, wy = ty.w;
  return a * ((1.0 - wx) * (1.0 - wy)) + b * (wx * (1.0 - wy)) + e * ((1.0 - wx) * wy) + d * (wx * wy);
}

// GetTransmittanceToTopAtmosphereBoundary (functions.glsl:473-480)
__device__ V3 transmittance_to_top(const Ctx& k, double r, double mu) {
  double x, y;
  transmittance_xy(k.g, r, mu, &x, &y);
  return fetch2(k.t.transmittance, k.g.sz.t_w, tap_d(x, k.g.sz.t_w), tap_d(y, k.g.sz.t_h));
}
__device__ __forceinline__ V3 min1(V3 a) { return v3(fmin(a.x, 1.0), fmin(a.y, 1.0), fmin(a.z, 1.0)); }
__device__ __forceinline__ V3 div3(V3 a, V3 b) { return v3(a.x / b.x, a.y / b.y, a.z / b.z); }
// GetTransmittance (functions.glsl:493-519)
__device__ V3 get_transmittance(const Ctx& k, double r, double mu, double d, bool hit) {
  const double r_d = d_clamp(sqrt(d * d + 2.0 * r * mu * d + r * r), k.g.bottom, k.g.top);
  const double mu_d = d_clamp((r * mu + d) / r_d, -1.0, 1.0);
  if (hit) {
    return min1(div3(transmittance_to_top(k, r_d, -mu_d), transmittance_to_top(k, r, -mu)));
  }
  return min1(div3(transmittance_to_top(k, r, mu), transmittance_to_top(k, r_d, mu_d)));
}
__device__ __forceinline__ double smoothstep_d(double e0, double e1, double x) {
  x = d_clamp((x - e0) / (e1 - e0), 0.0, 1.0);
  return x * x * (3.0 - 2.0 * x);
}
// GetTransmittanceToSun (functions.glsl:552-563)
__device__ V3 transmittance_to_sun(const Ctx& k, double r, double mu_s) {
  const double sin_h = k.g.bottom / r;
  const double cos_h = -sqrt(d_pos(1.0 - sin_h * sin_h));
  const double a = k.g.sun_angular_radius;
  return transmittance_to_top(k, r, mu_s) * smoothstep_d(-sin_h * a, sin_h * a, mu_s - cos_h);
}
// GetIrradiance (functions.glsl:1524-1533, 1595-1601)
__device__ V3 get_irradiance(const Ctx& k, double r, double mu_s) {
  const double x_r = (r - k.g.bottom) / (k.g.top - k.g.bottom);
  const double x_mu_s = mu_s * 0.5 + 0.5;
  return fetch2(k.t.irradiance, k.g.sz.e_w, tap_d(x_mu_s * (k.g.sz.e_w - 1), k.g.sz.e_w),
                tap_d(x_r * (k.g.sz.e_h - 1), k.g.sz.e_h));
}

// GetExtrapolatedSingleMieScattering (functions.glsl:1634-1646)
__device__ V3 extrapolate_single_mie(const Ctx& k, V3 scattering, double mie_red) {
  if (scattering.x <= 0.0) return splat(0.0);
  const double f = mie_red / scattering.x * (k.c.rayleigh[0] / k.c.mie_sca[0]);
  return v3(scattering.x * f * (k.c.mie_sca[0] / k.c.rayleigh[0]),
            scattering.y * f * (k.c.mie_sca[1] / k.c.rayleigh[1]),
            scattering.z * f * (k.c.mie_sca[2] / k.c.rayleigh[2]));
}

// GetCombinedScattering (functions.glsl:1658-1690): one shared 4-D footprint, 16 texels per table.
__device__ V3 combined_scattering(const Ctx& k, double r, double mu, double mu_s, double nu, bool hit,
                                  V3* single_mie) {
  const PasSizes& z = k.g.sz;
  // GetScatteringTextureUvwzFromRMuMuSNu (functions.glsl:773-831) in texel space
  const double rho = sqrt(d_pos(r * r - k.g.bottom * k.g.bottom));
  const TapD tz = tap_d(rho / k.g.H * (z.r_n - 1), z.r_n);
  const TapD ty = tap_d(scattering_y_from_mu(k.g, r, rho, mu, hit), z.mu_n);
  const double xs = scattering_x_from_mu_s(k.g, mu_s);
  const double tex_coord_x = (nu + 1.0) * 0.5 * (z.nu_n - 1);
  const double tex_x = floor(tex_coord_x);
  const double lerp = tex_coord_x - tex_x;
  const int width = z.nu_n * z.mu_s_n;
  const TapD tx[2] = {tap_d(tex_x * z.mu_s_n + xs, width), tap_d((tex_x + 1.0) * z.mu_s_n + xs, width)};
  const double wslab[2] = {1.0 - lerp, lerp};
  double s[4] = {0, 0, 0, 0}, m[3] = {0, 0, 0};
#pragma unroll
  for (int q = 0; q < 2; ++q) {
#pragma unroll
    for (int corner = 0; corner < 8; ++corner) {
      const int ix = (corner & 1) ? tx[q].i1 : tx[q].i0;
      const int iy = (corner & 2) ? ty.i1 : ty.i0;
      const int iz = (corner & 4) ? tz.i1 : tz.i0;
      const double w = wslab[q] * ((corner & 1) ? tx[q].w : 1.0 - tx[q].w) *
                       ((corner & 2) ? ty.w : 1.0 - ty.w) * ((corner & 4) ? tz.w : 1.0 - tz.w);
      const size_t texel = ix + (size_t)width * (iy + (size_t)z.mu_n * iz);
      const Rgba v = texel_rgba(k.t.scattering, k.t.half_precision, texel);
      s[0] += w * v.r; s[1] += w * v.g; s[2] += w * v.b; s[3] += w * v.a;
      if (k.t.single_mie != nullptr) {
        const Rgba u = texel_rgba(k.t.single_mie, k.t.half_precision, texel);
        m[0] += w * u.r; m[1] += w * u.g; m[2] += w * u.b;
      }
    }
  }
  const V3 scattering = v3(s[0], s[1], s[2]);
  *single_mie = k.t.single_mie != nullptr ? v3(m[0], m[1], m[2]) : extrapolate_single_mie(k, scattering, s[3]);
  return scattering;
}

__device__ __forceinline__ double mie_phase(double g, double nu) {
  return mie_phase_k(g) * (1.0 + nu * nu) / pow(1.0 + g * g - 2.0 * g * nu, 1.5);
}

// GetSkyRadiance (functions.glsl:1705-1769), times SKY_SPECTRAL_RADIANCE_TO_LUMINANCE in luminance mode
__device__ V3 sky_radiance(const Ctx& k, V3 camera, V3 view_ray, double shadow_length, V3 sun_direction,
                           V3* transmittance) {
  double r = sqrt(dot(camera, camera));
  double rmu = dot(camera, view_ray);
  const double dist_to_top = -rmu - sqrt(rmu * rmu - r * r + k.g.top * k.g.top);
  if (dist_to_top > 0.0) {
    camera = camera + view_ray * dist_to_top;
    r = k.g.top;
    rmu += dist_to_top;
  } else if (r > k.g.top) {
    *transmittance = splat(1.0);
    return splat(0.0);
  }
  const double mu = rmu / r;
  const double mu_s = dot(camera, sun_direction) / r;
  const double nu = dot(view_ray, sun_direction);
  const bool hit = hits_ground(k.g, r, mu);
  *transmittance = hit ? splat(0.0) : transmittance_to_top(k, r, mu);
  V3 single_mie, scattering;
  if (shadow_length == 0.0) {
    scattering = combined_scattering(k, r, mu, mu_s, nu, hit, &single_mie);
  } else {
    const double d = shadow_length;
    const double r_p = d_clamp(sqrt(d * d + 2.0 * r * mu * d + r * r), k.g.bottom, k.g.top);
    const double mu_p = (r * mu + d) / r_p;
    const double mu_s_p = (r * mu_s + d * nu) / r_p;
    scattering = combined_scattering(k, r_p, mu_p, mu_s_p, nu, hit, &single_mie);
    const V3 shadow_t = get_transmittance(k, r, mu, shadow_length, hit);
    scattering = scattering * shadow_t;
    single_mie = single_mie * shadow_t;
  }
  const V3 L = scattering * rayleigh_phase(nu) + single_mie * mie_phase(k.g.mie_g, nu);
  return L * load3(k.c.sky_k);
}

// GetSkyRadianceToPoint (functions.glsl:1787-1863)
__device__ V3 sky_radiance_to_point(const Ctx& k, V3 camera, V3 point, double shadow_length,
                                    V3 sun_direction, V3* transmittance) {
  V3 view_ray = point - camera;
  view_ray = view_ray * (1.0 / sqrt(dot(view_ray, view_ray)));
  double r = sqrt(dot(camera, camera));
  double rmu = dot(camera, view_ray);
  const double dist_to_top = -rmu - sqrt(rmu * rmu - r * r + k.g.top * k.g.top);
  if (dist_to_top > 0.0) {
    camera = camera + view_ray * dist_to_top;
    r = k.g.top;
    rmu += dist_to_top;
  }
  const double mu = rmu / r;
  const double mu_s = dot(camera, sun_direction) / r;
  const double nu = dot(view_ray, sun_direction);
  const V3 pc = point - camera;
  double d = sqrt(dot(pc, pc));
  const bool hit = hits_ground(k.g, r, mu);
  *transmittance = get_transmittance(k, r, mu, d, hit);
  V3 single_mie;
  V3 scattering = combined_scattering(k, r, mu, mu_s, nu, hit, &single_mie);
  d = fmax(d - shadow_length, 0.0);
  const double r_p = d_clamp(sqrt(d * d + 2.0 * r * mu * d + r * r), k.g.bottom, k.g.top);
  const double mu_p = (r * mu + d) / r_p;
  const double mu_s_p = (r * mu_s + d * nu) / r_p;
  V3 single_mie_p;
  const V3 scattering_p = combined_scattering(k, r_p, mu_p, mu_s_p, nu, hit, &single_mie_p);
  V3 shadow_t = *transmittance;
  if (shadow_length > 0.0) shadow_t = get_transmittance(k, r, mu, d, hit);
  scattering = scattering - shadow_t * scattering_p;
  single_mie = single_mie - shadow_t * single_mie_p;
  if (k.t.single_mie == nullptr) {
    single_mie = extrapolate_single_mie(k, scattering, single_mie.x);  // functions.glsl:1851-1854
  }
  single_mie = single_mie * smoothstep_d(0.0, 0.01, mu_s);
  const V3 L = scattering * rayleigh_phase(nu) + single_mie * mie_phase(k.g.mie_g, nu);
  return L * load3(k.c.sky_k);
}

// GetSunAndSkyIrradiance (functions.glsl:1878-1896) / GetSunAndSkyIlluminance (model.cc:272-280)
__device__ V3 sun_and_sky_irradiance(const Ctx& k, V3 point, V3 normal, V3 sun_direction, V3* sky_irradiance) {
  const double r = sqrt(dot(point, point));
  const double mu_s = dot(point, sun_direction) / r;
  *sky_irradiance = get_irradiance(k, r, mu_s) * ((1.0 + dot(normal, point) / r) * 0.5) * load3(k.c.sky_k);
  return load3(k.c.solar) * transmittance_to_sun(k, r, mu_s) * fmax(dot(normal, sun_direction), 0.0) *
         load3(k.c.sun_k);
}

// ---- the test scene (reference/model_test.glsl) ------------------------------------------------
__device__ double sun_visibility(const RenderView& s, V3 point, V3 sun_direction) {
  const V3 p = point - load3(s.sphere_center);
  const double p_dot_v = dot(p, sun_direction), p_dot_p = dot(p, p);
  const double d2 = p_dot_p - p_dot_v * p_dot_v;
  const double dist = -p_dot_v - sqrt(s.sphere_radius * s.sphere_radius - d2);
  if (dist > 0.0) {
    const double ray_sphere_distance = s.sphere_radius - sqrt(d2);
    return smoothstep_d(1.0, 0.0, (-ray_sphere_distance / p_dot_v) / s.sun_size[0]);
  }
  return 1.0;
}
__device__ double sky_visibility(const RenderView& s, V3 point) {
  const V3 p = point - load3(s.sphere_center);
  const double p_dot_p = dot(p, p);
  return 1.0 + p.z / sqrt(p_dot_p) * s.sphere_radius * s.sphere_radius / p_dot_p;
}
__device__ void sphere_shadow_in_out(const RenderView& s, V3 view_direction, double* d_in, double* d_out) {
  const V3 pos = load3(s.camera) - load3(s.sphere_center);
  const V3 sun = load3(s.sun_direction);
  const double pos_dot_sun = dot(pos, sun), view_dot_sun = dot(view_direction, sun);
  const double kk = s.sun_size[0], R = s.sphere_radius;
  const double l = 1.0 + kk * kk;
  const double a = 1.0 - l * view_dot_sun * view_dot_sun;
  const double b = dot(pos, view_direction) - l * pos_dot_sun * view_dot_sun - kk * R * view_dot_sun;
  const double c = dot(pos, pos) - l * pos_dot_sun * pos_dot_sun - 2.0 * kk * R * pos_dot_sun - R * R;
  const double disc = b * b - a * c;
  if (disc > 0.0) {
    *d_in = fmax(0.0, (-b - sqrt(disc)) / a);
    *d_out = (-b + sqrt(disc)) / a;
    const double d_base = -pos_dot_sun / view_dot_sun;
    const double d_apex = -(pos_dot_sun + R / kk) / view_dot_sun;
    if (view_dot_sun > 0.0) {
      *d_in = fmax(*d_in, d_apex);
      *d_out = a > 0.0 ? fmin(*d_out, d_base) : d_base;
    } else {
      *d_in = a > 0.0 ? fmax(*d_in, d_base) : d_base;
      *d_out = fmin(*d_out, d_apex);
    }
  } else {
    *d_in = 0.0;
    *d_out = 0.0;
  }
}

// GetViewRayRadiance (model_test.glsl:218-348)
__device__ V3 view_ray_radiance(const Ctx& k, const RenderView& s, V3 view_ray, V3 view_ray_diff) {
  const double vlen = sqrt(dot(view_ray, view_ray));
  const V3 v = view_ray * (1.0 / vlen);
  const double fragment_angular_size = sqrt(dot(view_ray_diff, view_ray_diff)) / vlen;
  double shadow_in, shadow_out;
  sphere_shadow_in_out(s, v, &shadow_in, &shadow_out);
  const V3 camera = load3(s.camera), earth_center = load3(s.earth_center), sun = load3(s.sun_direction);
  const V3 cam_e = camera - earth_center;
  const double inv_pi = 1.0 / kPi;

  V3 p = camera - load3(s.sphere_center);
  double p_dot_v = dot(p, v), p_dot_p = dot(p, p);
  double d2 = p_dot_p - p_dot_v * p_dot_v;
  double dist = -p_dot_v - sqrt(s.sphere_radius * s.sphere_radius - d2);
  double sphere_alpha = 0.0;
  V3 sphere_radiance = splat(0.0);
  if (dist > 0.0) {
    const double ray_sphere_distance = s.sphere_radius - sqrt(d2);
    sphere_alpha = fmin((-ray_sphere_distance / p_dot_v) / fragment_angular_size, 1.0);
    const V3 point = camera + v * dist;
    V3 normal = point - load3(s.sphere_center);
    normal = normal * (1.0 / sqrt(dot(normal, normal)));
    V3 sky_e;
    const V3 sun_e = sun_and_sky_irradiance(k, point - earth_center, normal, sun, &sky_e);
    sphere_radiance = load3(s.sphere_albedo) * inv_pi * (sun_e + sky_e);
    const double shadow_length = fmax(0.0, fmin(shadow_out, dist) - shadow_in);
    V3 tr;
    const V3 in_scatter = sky_radiance_to_point(k, cam_e, point - earth_center, shadow_length, sun, &tr);
    sphere_radiance = sphere_radiance * tr + in_scatter;
  }

  p_dot_v = dot(cam_e, v);
  p_dot_p = dot(cam_e, cam_e);
  d2 = p_dot_p - p_dot_v * p_dot_v;
  dist = -p_dot_v - sqrt(earth_center.z * earth_center.z - d2);
  double ground_alpha = 0.0;
  V3 ground_radiance = splat(0.0);
  if (dist > 0.0) {
    const V3 point = camera + v * dist;
    V3 normal = point - earth_center;
    normal = normal * (1.0 / sqrt(dot(normal, normal)));
    V3 sky_e;
    const V3 sun_e = sun_and_sky_irradiance(k, point - earth_center, normal, sun, &sky_e);
    ground_radiance = load3(s.ground_albedo) * inv_pi *
                      (sun_e * sun_visibility(s, point, sun) + sky_e * sky_visibility(s, point));
    const double shadow_length = fmax(0.0, fmin(shadow_out, dist) - shadow_in);
    V3 tr;
    const V3 in_scatter = sky_radiance_to_point(k, cam_e, point - earth_center, shadow_length, sun, &tr);
    ground_radiance = ground_radiance * tr + in_scatter;
    ground_alpha = 1.0;
  }

  const double shadow_length = fmax(0.0, shadow_out - shadow_in);
  V3 tr;
  V3 radiance = sky_radiance(k, cam_e, v, shadow_length, sun, &tr);
  if (dot(v, sun) > s.sun_size[1]) {
    // GetSolarRadiance / GetSolarLuminance (model.cc:228-231, 254-258)
    const double a = k.g.sun_angular_radius;
    radiance = radiance + tr * (load3(k.c.solar) * (1.0 / (kPi * a * a)) * load3(k.c.sun_k));
  }
  radiance = radiance * (1.0 - ground_alpha) + ground_radiance * ground_alpha;
  radiance = radiance * (1.0 - sphere_alpha) + sphere_radiance * sphere_alpha;
  return radiance;
}

__device__ __forceinline__ unsigned tone(double v, double exposure) {
  // model_test.cc:726-731: pow(1 - exp(-v * exposure), 1 / 2.2), truncated to 8 bits
  const double t = pow(1.0 - exp(-v * exposure), 1.0 / 2.2);
  return (unsigned)(t * 255.0);
}

__global__ void __launch_bounds__(128)
render_scene_kernel(const __grid_constant__ PasGeometry g, const __grid_constant__ RenderTables t,
                    const __grid_constant__ RenderConstants c, const __grid_constant__ RenderView s,
                    float* __restrict__ rgb, unsigned* __restrict__ argb) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y * blockDim.y + threadIdx.y;
  if (i >= s.width || j >= s.height) return;
  const Ctx k{g, t, c};
  // view rays of pixel (i, j), j = 0 at the top (model_test.cc:690-711)
  const double y = 1.0 - 2.0 * (j + 0.5) / s.height, dy = -2.0 / s.height;
  const double x = 2.0 * (i + 0.5) / s.width - 1.0, dx = 2.0 / s.width;
  const double* M = s.model_from_clip;
  const V3 view_ray = v3(M[0] * x + M[1] * y + M[2], M[3] * x + M[4] * y + M[5], M[6] * x + M[7] * y + M[8]);
  const V3 diff = v3(M[0] * dx + M[1] * dy, M[3] * dx + M[4] * dy, M[6] * dx + M[7] * dy);
  const V3 L = view_ray_radiance(k, s, view_ray, diff);
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

// Batched point queries: vectors are [n][3] doubles, outputs [n][3] floats.
__global__ void __launch_bounds__(128)
sky_radiance_kernel(const __grid_constant__ PasGeometry g, const __grid_constant__ RenderTables t,
                    const __grid_constant__ RenderConstants c, size_t n, int to_point,
                    const double* __restrict__ camera, const double* __restrict__ target,
                    const double* __restrict__ shadow_length, const double* __restrict__ sun_direction,
                    float* __restrict__ radiance, float* __restrict__ transmittance) {
  const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n) return;
  const Ctx k{g, t, c};
  V3 tr;
  const double sl = shadow_length != nullptr ? shadow_length[q] : 0.0;
  const V3 L = to_point
                   ? sky_radiance_to_point(k, load3(camera + 3 * q), load3(target + 3 * q), sl,
                                           load3(sun_direction + 3 * q), &tr)
                   : sky_radiance(k, load3(camera + 3 * q), load3(target + 3 * q), sl,
                                  load3(sun_direction + 3 * q), &tr);
  radiance[3 * q + 0] = (float)L.x; radiance[3 * q + 1] = (float)L.y; radiance[3 * q + 2] = (float)L.z;
  if (transmittance != nullptr) {
    transmittance[3 * q + 0] = (float)tr.x; transmittance[3 * q + 1] = (float)tr.y; transmittance[3 * q + 2] = (float)tr.z;
  }
}

__global__ void __launch_bounds__(128)
sun_and_sky_irradiance_kernel(const __grid_constant__ PasGeometry g, const __grid_constant__ RenderTables t,
                              const __grid_constant__ RenderConstants c, size_t n,
                              const double* __restrict__ point, const double* __restrict__ normal,
                              const double* __restrict__ sun_direction, float* __restrict__ sun_irradiance,
                              float* __restrict__ sky_irradiance) {
  const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n) return;
  const Ctx k{g, t, c};
  V3 sky;
  const V3 sun = sun_and_sky_irradiance(k, load3(point + 3 * q), load3(normal + 3 * q),
                                        load3(sun_direction + 3 * q), &sky);
  sun_irradiance[3 * q + 0] = (float)sun.x; sun_irradiance[3 * q + 1] = (float)sun.y; sun_irradiance[3 * q + 2] = (float)sun.z;
  sky_irradiance[3 * q + 0] = (float)sky.x; sky_irradiance[3 * q + 1] = (float)sky.y; sky_irradiance[3 * q + 2] = (float)sky.z;
}

}  // namespace

cudaError_t launch_render_scene(const PasGeometry& g, const RenderTables& t, const RenderConstants& c,
                                const RenderView& view, float* rgb, unsigned* argb, cudaStream_t stream) {
  const dim3 block(32, 4);
  const dim3 grid((view.width + block.x - 1) / block.x, (view.height + block.y - 1) / block.y);
  render_scene_kernel<<<grid, block, 0, stream>>>(g, t, c, view, rgb, argb);
  return cudaGetLastError();
}

cudaError_t launch_sky_radiance(const PasGeometry& g, const RenderTables& t, const RenderConstants& c,
                                size_t n, bool to_point, const double* camera, const double* target,
                                const double* shadow_length, const double* sun_direction, float* radiance,
                                float* transmittance, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  sky_radiance_kernel<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(
      g, t, c, n, to_point ? 1 : 0, camera, target, shadow_length, sun_direction, radiance, transmittance);
  return cudaGetLastError();
}

cudaError_t launch_sun_and_sky_irradiance(const PasGeometry& g, const RenderTables& t,
                                          const RenderConstants& c, size_t n, const double* point,
                                          const double* normal, const double* sun_direction,
                                          float* sun_irradiance, float* sky_irradiance, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  sun_and_sky_irradiance_kernel<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(
      g, t, c, n, point, normal, sun_direction, sun_irradiance, sky_irradiance);
  return cudaGetLastError();
}

}  // namespace pas

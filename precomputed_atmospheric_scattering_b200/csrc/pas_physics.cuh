// Host/device inline physics shared by every kernel of the engine and by the host code of
// pas_model.cu. The header also compiles with a plain C++ compiler: tests/emu/physics_host.cc
// exports it to tests/test_physics_header.py, which checks every mapping against the oracle on the CPU.
//
// Two tiers:
//   * fp64 "setup" math: texel -> (r, mu, mu_s, nu) inverse mappings and the per-(layer,
//     direction) / per-sample geometry that is shared by many threads. Runs once per texel or per
//     block, so double precision costs nothing and removes the catastrophic cancellations of the
//     as-written expressions (atmosphere/functions.glsl:211-212, 411, 792-793).
//   * fp32 "inner loop" math: what runs per direction / per ray sample per thread. Written in
//     cancellation-free forms (altitudes instead of radii, rationalised quadratic roots) and
//     using texel-space coordinates directly: for a table axis of n texels the reference's
//     u = 0.5/n + x(1 - 1/n) followed by the fetch rule u*n - 0.5
//     (functions.glsl:342-343; dimensional_types/math/binary_function.h:103-113) is simply
//     x * (n - 1).
#ifndef PAS_B200_CSRC_PAS_PHYSICS_CUH_
#define PAS_B200_CSRC_PAS_PHYSICS_CUH_

#include <math.h>

#include "pas_types.h"

#if defined(__CUDACC__)
#define PAS_HD __host__ __device__ __forceinline__
#else
#define PAS_HD inline
#endif

namespace pas {

constexpr double kPi = 3.14159265358979323846;

// ---- portable fp32 intrinsics ------------------------------------------------------------------
PAS_HD float f_rsqrt(float x) {
#if defined(__CUDA_ARCH__)
  return rsqrtf(x);
#else
  return 1.0f / sqrtf(x);
#endif
}
// Reciprocal and square root of the fp32 inner loops go straight to the SFU (MUFU.RCP / MUFU.SQRT,
// <= 2 ulp, no denormal fix-up branches); the arguments are table coordinates and distances whose
// 1e-7 relative error is far below the 1e-3 parity budget.
PAS_HD float f_rcp(float x) {
#if defined(__CUDA_ARCH__)
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
#else
  return 1.0f / x;
#endif
}
PAS_HD float f_sqrt(float x) {
#if defined(__CUDA_ARCH__)
  float y;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
#else
  return sqrtf(x);
#endif
}
PAS_HD float f_sat(float x) {
#if defined(__CUDA_ARCH__)
  return __saturatef(x);
#else
  return x < 0.0f ? 0.0f : (x > 1.0f ? 1.0f : x);
#endif
}
PAS_HD float f_clamp(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
PAS_HD double d_clamp(double x, double lo, double hi) { return fmin(fmax(x, lo), hi); }
PAS_HD double d_pos(double x) { return fmax(x, 0.0); }

// ---- fp64 setup tier ---------------------------------------------------------------------------
// GetUnitRangeFromTextureCoord / GetTextureCoordFromUnitRange (functions.glsl:342-348).
PAS_HD double unit_from_coord(double u, int n) { return (u - 0.5 / n) / (1.0 - 1.0 / n); }
PAS_HD double coord_from_unit(double x, int n) { return 0.5 / n + x * (1.0 - 1.0 / n); }

// Density profile (functions.glsl:263-273).
PAS_HD double profile_density(const double (*P)[5], double h) {
  const double* L = h < P[0][0] ? P[0] : P[1];
  // (a layer without exponential term -- the ozone layers, the zero layers padded in front -- skips the
  // fp64 exp: 0 * exp(x) is exactly 0 for every finite exp(x))
  const double e = L[1] != 0.0 ? L[1] * exp(L[2] * h) : 0.0;
  return d_clamp(e + L[3] * h + L[4], 0.0, 1.0);
}

// Distances to the boundaries and the ground test (functions.glsl:207-246), as written.
PAS_HD double dist_top(const PasGeometry& g, double r, double mu) {
  return d_pos(-r * mu + sqrt(d_pos(r * r * (mu * mu - 1.0) + g.top * g.top)));
}
PAS_HD double dist_bottom(const PasGeometry& g, double r, double mu) {
  return d_pos(-r * mu - sqrt(d_pos(r * r * (mu * mu - 1.0) + g.bottom * g.bottom)));
}
PAS_HD bool hits_ground(const PasGeometry& g, double r, double mu) {
  return mu < 0.0 && r * r * (mu * mu - 1.0) + g.bottom * g.bottom >= 0.0;
}

// One table axis: texel-space coordinate -> (i0, i1, w1) with the reference's clamping rule.
struct Tap {
  int i0, i1;
  float w;  // weight of i1
};
PAS_HD Tap make_tap(double x, int n) {  // x in texel space (u*n - 0.5)
  Tap t;
  double fl = floor(x);
  int i = (int)fl;
  t.w = (float)(x - fl);
  t.i0 = i < 0 ? 0 : (i > n - 1 ? n - 1 : i);
  t.i1 = i + 1 < 0 ? 0 : (i + 1 > n - 1 ? n - 1 : i + 1);
  return t;
}
PAS_HD Tap make_tap_f(float x, int n) {  // fp32 variant; x already clamped to [0, n-1]
  Tap t;
  float fl = floorf(x);
  int i = (int)fl;
  i = i > n - 2 ? n - 2 : i;
  i = i < 0 ? 0 : i;
  t.i0 = i;
  t.i1 = i + 1 > n - 1 ? n - 1 : i + 1;
  t.w = x - (float)i;
  return t;
}

// r of scattering layer k / transmittance row / irradiance row (functions.glsl:437-438, 849-851,
// 1545-1546). rho is returned too because r^2 - bottom^2 == rho^2 exactly by construction.
PAS_HD void layer_radius(const PasGeometry& g, double u_r, int n, double* r, double* rho) {
  *rho = g.H * unit_from_coord(u_r, n);
  *r = sqrt(*rho * *rho + g.bottom * g.bottom);
}

// Transmittance-table texel-space coordinates of (r, mu) (functions.glsl:402-421).
PAS_HD void transmittance_xy(const PasGeometry& g, double r, double mu, double* x, double* y) {
  double rho = sqrt(d_pos(r * r - g.bottom * g.bottom));
  double d = dist_top(g, r, mu);
  double d_min = g.top - r, d_max = rho + g.H;
  *x = (d - d_min) / (d_max - d_min) * (g.sz.t_w - 1);
  *y = rho / g.H * (g.sz.t_h - 1);
}

// mu of scattering row j at radius (r, rho) (functions.glsl:853-875) and whether the ray hits the
// ground (rows of the lower half). Also returns r*mu computed without dividing by r.
PAS_HD void scattering_row_mu(const PasGeometry& g, double r, double rho, int j, double* mu,
                              double* r_mu, bool* hit) {
  const int mu_n = g.sz.mu_n;
  double u_mu = (j + 0.5) / mu_n;
  if (u_mu < 0.5) {
    double d_min = r - g.bottom, d_max = rho;
    double d = d_min + (d_max - d_min) * unit_from_coord(1.0 - 2.0 * u_mu, mu_n / 2);
    *mu = d == 0.0 ? -1.0 : d_clamp(-(rho * rho + d * d) / (2.0 * r * d), -1.0, 1.0);
    *hit = true;
  } else {
    double d_min = g.top - r, d_max = rho + g.H;
    double d = d_min + (d_max - d_min) * unit_from_coord(2.0 * u_mu - 1.0, mu_n / 2);
    *mu = d == 0.0 ? 1.0
                   : d_clamp((g.H * g.H - rho * rho - d * d) / (2.0 * r * d), -1.0, 1.0);
    *hit = false;
  }
  *r_mu = r * *mu;
}

// mu_s of scattering column i_mu_s (functions.glsl:877-887).
PAS_HD double scattering_col_mu_s(const PasGeometry& g, int i_mu_s) {
  double x = unit_from_coord((i_mu_s + 0.5) / g.sz.mu_s_n, g.sz.mu_s_n);
  double d_min = g.top - g.bottom, d_max = g.H;
  double A = g.mus_A;
  double a = (A - x * A) / (1.0 + x * A);
  double d = d_min + fmin(a, A) * (d_max - d_min);
  return d == 0.0 ? 1.0 : d_clamp((g.H * g.H - d * d) / (2.0 * g.bottom * d), -1.0, 1.0);
}

// nu of slab i_nu, clamped to the range allowed by (mu, mu_s) (functions.glsl:889, 923-925).
PAS_HD double scattering_slab_nu(const PasGeometry& g, int i_nu, double mu, double mu_s) {
  double nu = d_clamp((double)i_nu / (g.sz.nu_n - 1) * 2.0 - 1.0, -1.0, 1.0);
  double s = sqrt((1.0 - mu * mu) * (1.0 - mu_s * mu_s));
  return d_clamp(nu, mu * mu_s - s, mu * mu_s + s);
}

// Forward mu mapping of the scattering table in texel space (functions.glsl:789-812): returns
// u_mu * mu_n - 0.5 for the ray (r, mu) with the given ground flag.
PAS_HD double scattering_y_from_mu(const PasGeometry& g, double r, double rho, double mu, bool hit) {
  const int half = g.sz.mu_n / 2;
  double r_mu = r * mu;
  double disc = r_mu * r_mu - r * r + g.bottom * g.bottom;
  double u_mu;
  if (hit) {
    double d = -r_mu - sqrt(d_pos(disc));
    double d_min = r - g.bottom, d_max = rho;
    u_mu = 0.5 - 0.5 * coord_from_unit(d_max == d_min ? 0.0 : (d - d_min) / (d_max - d_min), half);
  } else {
    double d = -r_mu + sqrt(d_pos(disc + g.H * g.H));
    double d_min = g.top - r, d_max = rho + g.H;
    u_mu = 0.5 + 0.5 * coord_from_unit((d - d_min) / (d_max - d_min), half);
  }
  return u_mu * g.sz.mu_n - 0.5;
}

// Forward mu_s mapping in texel space (functions.glsl:814-827): u_mu_s * mu_s_n - 0.5.
PAS_HD double scattering_x_from_mu_s(const PasGeometry& g, double mu_s) {
  double d = dist_top(g, g.bottom, mu_s);
  double d_min = g.top - g.bottom, d_max = g.H;
  double a = (d - d_min) / (d_max - d_min);
  return d_pos(1.0 - a / g.mus_A) / (1.0 + a) * (g.sz.mu_s_n - 1);
}

// Phase functions (functions.glsl:739-747).
PAS_HD double rayleigh_phase(double nu) { return 3.0 / (16.0 * kPi) * (1.0 + nu * nu); }
PAS_HD double mie_phase_k(double g) { return 3.0 / (8.0 * kPi) * (1.0 - g * g) / (2.0 + g * g); }

// ---- fp32 inner-loop tier ----------------------------------------------------------------------
// Distance to the top boundary from a point given p = r*mu and q = (top - r)(top + r) >= 0,
// rationalised so that neither sign of p cancels (functions.glsl:207-214 restated).
PAS_HD float f_dist_top(float p, float q) {
  float s = f_sqrt(fmaf(p, p, q));
  return p > 0.0f ? q * f_rcp(p + s) : s - p;
}

// mu_s axis of the scattering table, texel space, from p = bottom * mu_s (functions.glsl:814-827):
// d = dist_top(bottom, mu_s); a = (d - d_min) / (d_max - d_min); x = max(1 - a/A, 0) / (1 + a).
struct MuSMap {
  float H2;         // top^2 - bottom^2
  float d_min;      // top - bottom
  float inv_range;  // 1 / (H - d_min)
  float inv_A;
  float scale;      // mu_s_n - 1
};
PAS_HD float f_mu_s_texel_x(const MuSMap& m, float p) {
  float d = f_dist_top(p, m.H2);
  float a = (d - m.d_min) * m.inv_range;
  return fmaxf(fmaf(-a, m.inv_A, 1.0f), 0.0f) * f_rcp(1.0f + a) * m.scale;
}

}  // namespace pas

#endif  // PAS_B200_CSRC_PAS_PHYSICS_CUH_

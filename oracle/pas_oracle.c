/* TEST INFRASTRUCTURE ONLY -- see pas_oracle.h. Double-precision CPU restatement of the LUT
 * precomputation hot path. Every function cites the reference lines it restates; paths are
 * relative to the reference root (ebruneton/precomputed_atmospheric_scattering @ d9954923).
 * "f.glsl" abbreviates atmosphere/functions.glsl.
 *
 * Conventions: doubles throughout; spectra are arrays of a->nc channels; geometry (anything that
 * does not depend on wavelength) is computed once and shared by the channel loops.
 */
#include "pas_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdatomic.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#define PI_D 3.14159265358979323846

typedef paso_atmosphere Atm;

/* ---- scalar helpers (f.glsl:113-127) ------------------------------------------------------- */
static inline double clampd(double x, double lo, double hi) {
  return x < lo ? lo : (x > hi ? hi : x);
}
static inline double clamp_cos(double mu) { return clampd(mu, -1.0, 1.0); }
static inline double clamp_radius(const Atm* a, double r) {
  return clampd(r, a->bottom_radius, a->top_radius);
}
static inline double safe_sqrt(double x) { return sqrt(x > 0.0 ? x : 0.0); }
static inline double pos(double x) { return x > 0.0 ? x : 0.0; }

/* ---- ray / sphere intersections (f.glsl:207-246) ------------------------------------------- */
double paso_distance_to_top(const Atm* a, double r, double mu) {
  double disc = r * r * (mu * mu - 1.0) + a->top_radius * a->top_radius;
  return pos(-r * mu + safe_sqrt(disc));
}
double paso_distance_to_bottom(const Atm* a, double r, double mu) {
  double disc = r * r * (mu * mu - 1.0) + a->bottom_radius * a->bottom_radius;
  return pos(-r * mu - safe_sqrt(disc));
}
int paso_ray_intersects_ground(const Atm* a, double r, double mu) {
  return mu < 0.0 &&
         r * r * (mu * mu - 1.0) + a->bottom_radius * a->bottom_radius >= 0.0;
}

/* ---- density profiles (f.glsl:263-273) ----------------------------------------------------- */
static double layer_density(const double* L, double h) {
  double d = L[1] * exp(L[2] * h) + L[3] * h + L[4];
  return clampd(d, 0.0, 1.0);
}
double paso_profile_density(const Atm* a, int profile, double h) {
  const double(*P)[5] = a->profiles[profile];
  return h < P[0][0] ? layer_density(P[0], h) : layer_density(P[1], h);
}

/* ---- optical length, 500-interval trapezoid (f.glsl:275-299) ------------------------------- */
double paso_optical_length_to_top(const Atm* a, int profile, double r, double mu) {
  const int n = 500;
  double dx = paso_distance_to_top(a, r, mu) / n;
  double sum = 0.0;
  for (int i = 0; i <= n; ++i) {
    double d = i * dx;
    double ri = sqrt(d * d + 2.0 * r * mu * d + r * r);
    double y = paso_profile_density(a, profile, ri - a->bottom_radius);
    sum += y * ((i == 0 || i == n) ? 0.5 : 1.0) * dx;
  }
  return sum;
}

/* ---- transmittance to the top boundary (f.glsl:306-320) ------------------------------------ */
void paso_compute_transmittance_to_top(const Atm* a, double r, double mu, double* out) {
  double tr = paso_optical_length_to_top(a, 0, r, mu);
  double tm = paso_optical_length_to_top(a, 1, r, mu);
  double ta = paso_optical_length_to_top(a, 2, r, mu);
  for (int c = 0; c < a->nc; ++c) {
    out[c] = exp(-(a->rayleigh_scattering[c] * tr + a->mie_extinction[c] * tm +
                   a->absorption_extinction[c] * ta));
  }
}

/* ---- unit range <-> texture coordinate (f.glsl:342-348) ------------------------------------ */
static inline double coord_from_unit(double x, int n) { return 0.5 / n + x * (1.0 - 1.0 / n); }
static inline double unit_from_coord(double u, int n) { return (u - 0.5 / n) / (1.0 - 1.0 / n); }

/* ---- transmittance table parameterisation (f.glsl:402-447) --------------------------------- */
void paso_transmittance_uv_from_rmu(const Atm* a, double r, double mu, double* uv) {
  double H = sqrt(a->top_radius * a->top_radius - a->bottom_radius * a->bottom_radius);
  double rho = safe_sqrt(r * r - a->bottom_radius * a->bottom_radius);
  double d = paso_distance_to_top(a, r, mu);
  double d_min = a->top_radius - r, d_max = rho + H;
  uv[0] = coord_from_unit((d - d_min) / (d_max - d_min), a->sz.t_w);
  uv[1] = coord_from_unit(rho / H, a->sz.t_h);
}
void paso_rmu_from_transmittance_uv(const Atm* a, double u, double v, double* rmu) {
  double x_mu = unit_from_coord(u, a->sz.t_w), x_r = unit_from_coord(v, a->sz.t_h);
  double H = sqrt(a->top_radius * a->top_radius - a->bottom_radius * a->bottom_radius);
  double rho = H * x_r;
  double r = sqrt(rho * rho + a->bottom_radius * a->bottom_radius);
  double d_min = a->top_radius - r, d_max = rho + H;
  double d = d_min + x_mu * (d_max - d_min);
  double mu = d == 0.0 ? 1.0 : (H * H - rho * rho - d * d) / (2.0 * r * d);
  rmu[0] = r;
  rmu[1] = clamp_cos(mu);
}

/* ---- software texture fetches ---------------------------------------------------------------
 * Exactly the CPU reference's rule: u = x*N - 0.5, floor, fractional weights, both indices
 * clamped to [0, N-1] (external/dimensional_types/math/binary_function.h:103-118 and
 * ternary_function.h:100-125). */
typedef struct { int i0, i1; double w0, w1; } Tap;
static inline Tap tap(double x, int n) {
  Tap t;
  double u = x * n - 0.5;
  int i = (int)floor(u);
  u -= i;
  t.i0 = i < 0 ? 0 : (i > n - 1 ? n - 1 : i);
  t.i1 = i + 1 < 0 ? 0 : (i + 1 > n - 1 ? n - 1 : i + 1);
  t.w0 = 1.0 - u;
  t.w1 = u;
  return t;
}
static void fetch2(const double* tab, int nx, int ny, int nc, double x, double y, double* out) {
  Tap a = tap(x, nx), b = tap(y, ny);
  size_t plane = (size_t)nx * ny;
  for (int c = 0; c < nc; ++c) {
    const double* p = tab + c * plane;
    out[c] = p[a.i0 + (size_t)nx * b.i0] * (a.w0 * b.w0) + p[a.i1 + (size_t)nx * b.i0] * (a.w1 * b.w0) +
             p[a.i0 + (size_t)nx * b.i1] * (a.w0 * b.w1) + p[a.i1 + (size_t)nx * b.i1] * (a.w1 * b.w1);
  }
}
static void fetch3(const double* tab, int nx, int ny, int nz, int nc, double x, double y, double z,
                   double scale, int accumulate, double* out) {
  Tap a = tap(x, nx), b = tap(y, ny), d = tap(z, nz);
  size_t plane = (size_t)nx * ny * nz;
  for (int c = 0; c < nc; ++c) {
    const double* p = tab + c * plane;
#define AT(i, j, k) p[(i) + (size_t)nx * ((j) + (size_t)ny * (k))]
    double v = AT(a.i0, b.i0, d.i0) * (a.w0 * b.w0 * d.w0) + AT(a.i1, b.i0, d.i0) * (a.w1 * b.w0 * d.w0) +
               AT(a.i0, b.i1, d.i0) * (a.w0 * b.w1 * d.w0) + AT(a.i1, b.i1, d.i0) * (a.w1 * b.w1 * d.w0) +
               AT(a.i0, b.i0, d.i1) * (a.w0 * b.w0 * d.w1) + AT(a.i1, b.i0, d.i1) * (a.w1 * b.w0 * d.w1) +
               AT(a.i0, b.i1, d.i1) * (a.w0 * b.w1 * d.w1) + AT(a.i1, b.i1, d.i1) * (a.w1 * b.w1 * d.w1);
#undef AT
    out[c] = accumulate ? out[c] + v * scale : v * scale;
  }
}

/* ---- transmittance lookups (f.glsl:473-563) ------------------------------------------------ */
static void transmittance_to_top(const Atm* a, const double* T, double r, double mu, double* out) {
  double uv[2];
  paso_transmittance_uv_from_rmu(a, r, mu, uv);
  fetch2(T, a->sz.t_w, a->sz.t_h, a->nc, uv[0], uv[1], out);
}
void paso_get_transmittance_to_top(const Atm* a, const double* T, double r, double mu, double* out) {
  transmittance_to_top(a, T, r, mu, out);
}
void paso_get_transmittance(const Atm* a, const double* T, double r, double mu, double d, int hit,
                            double* out) {
  double num[PASO_MAX_CHANNELS], den[PASO_MAX_CHANNELS];
  double r_d = clamp_radius(a, sqrt(d * d + 2.0 * r * mu * d + r * r));
  double mu_d = clamp_cos((r * mu + d) / r_d);
  if (hit) {
    transmittance_to_top(a, T, r_d, -mu_d, num);
    transmittance_to_top(a, T, r, -mu, den);
  } else {
    transmittance_to_top(a, T, r, mu, num);
    transmittance_to_top(a, T, r_d, mu_d, den);
  }
  for (int c = 0; c < a->nc; ++c) {
    double t = num[c] / den[c];
    out[c] = t < 1.0 ? t : 1.0;
  }
}
static inline double smoothstep_d(double e0, double e1, double x) {
  /* external/dimensional_types/math/scalar.h:296-300 */
  x = clampd((x - e0) / (e1 - e0), 0.0, 1.0);
  return x * x * (3.0 - 2.0 * x);
}
void paso_get_transmittance_to_sun(const Atm* a, const double* T, double r, double mu_s,
                                   double* out) {
  double sin_h = a->bottom_radius / r;
  double cos_h = -sqrt(pos(1.0 - sin_h * sin_h));
  double vis = smoothstep_d(-sin_h * a->sun_angular_radius, sin_h * a->sun_angular_radius,
                            mu_s - cos_h);
  transmittance_to_top(a, T, r, mu_s, out);
  for (int c = 0; c < a->nc; ++c) out[c] *= vis;
}

/* ---- single scattering (f.glsl:650-730) ---------------------------------------------------- */
static double distance_to_nearest_boundary(const Atm* a, double r, double mu, int hit) {
  return hit ? paso_distance_to_bottom(a, r, mu) : paso_distance_to_top(a, r, mu);
}
void paso_single_scattering_point(const Atm* a, const double* T, double r, double mu, double mu_s,
                                  double nu, int hit, double* rayleigh, double* mie) {
  const int n = 50;
  double dx = distance_to_nearest_boundary(a, r, mu, hit) / n;
  double sr[PASO_MAX_CHANNELS] = {0}, sm[PASO_MAX_CHANNELS] = {0};
  double t_path[PASO_MAX_CHANNELS], t_sun[PASO_MAX_CHANNELS];
  for (int i = 0; i <= n; ++i) {
    double d = i * dx;
    /* integrand, f.glsl:650-668 */
    double r_d = clamp_radius(a, sqrt(d * d + 2.0 * r * mu * d + r * r));
    double mu_s_d = clamp_cos((r * mu_s + d * nu) / r_d);
    paso_get_transmittance(a, T, r, mu, d, hit, t_path);
    paso_get_transmittance_to_sun(a, T, r_d, mu_s_d, t_sun);
    double dens_r = paso_profile_density(a, 0, r_d - a->bottom_radius);
    double dens_m = paso_profile_density(a, 1, r_d - a->bottom_radius);
    double w = (i == 0 || i == n) ? 0.5 : 1.0;
    for (int c = 0; c < a->nc; ++c) {
      double t = t_path[c] * t_sun[c];
      sr[c] += t * dens_r * w;
      sm[c] += t * dens_m * w;
    }
  }
  for (int c = 0; c < a->nc; ++c) {
    rayleigh[c] = sr[c] * dx * a->solar_irradiance[c] * a->rayleigh_scattering[c];
    mie[c] = sm[c] * dx * a->solar_irradiance[c] * a->mie_scattering[c];
  }
}

/* ---- phase functions (f.glsl:739-747) ------------------------------------------------------ */
double paso_rayleigh_phase(double nu) { return 3.0 / (16.0 * PI_D) * (1.0 + nu * nu); }
double paso_mie_phase(double g, double nu) {
  double k = 3.0 / (8.0 * PI_D) * (1.0 - g * g) / (2.0 + g * g);
  return k * (1.0 + nu * nu) / pow(1.0 + g * g - 2.0 * g * nu, 1.5);
}

/* ---- 4-D scattering table parameterisation (f.glsl:773-926) -------------------------------- */
void paso_scattering_uvwz_from_rmumusnu(const Atm* a, double r, double mu, double mu_s, double nu,
                                        int hit, double* uvwz) {
  const double b = a->bottom_radius, top = a->top_radius;
  double H = sqrt(top * top - b * b);
  double rho = safe_sqrt(r * r - b * b);
  double u_r = coord_from_unit(rho / H, a->sz.r);
  double r_mu = r * mu;
  double disc = r_mu * r_mu - r * r + b * b;
  double u_mu;
  if (hit) {
    double d = -r_mu - safe_sqrt(disc);
    double d_min = r - b, d_max = rho;
    u_mu = 0.5 - 0.5 * coord_from_unit(d_max == d_min ? 0.0 : (d - d_min) / (d_max - d_min),
                                       a->sz.mu / 2);
  } else {
    double d = -r_mu + safe_sqrt(disc + H * H);
    double d_min = top - r, d_max = rho + H;
    u_mu = 0.5 + 0.5 * coord_from_unit((d - d_min) / (d_max - d_min), a->sz.mu / 2);
  }
  double d = paso_distance_to_top(a, b, mu_s);
  double d_min = top - b, d_max = H;
  double aa = (d - d_min) / (d_max - d_min);
  double D = paso_distance_to_top(a, b, a->mu_s_min);
  double A = (D - d_min) / (d_max - d_min);
  double u_mu_s = coord_from_unit(pos(1.0 - aa / A) / (1.0 + aa), a->sz.mu_s);
  uvwz[0] = (nu + 1.0) / 2.0;
  uvwz[1] = u_mu_s;
  uvwz[2] = u_mu;
  uvwz[3] = u_r;
}
void paso_rmumusnu_from_scattering_uvwz(const Atm* a, const double* uvwz, double* out) {
  const double b = a->bottom_radius, top = a->top_radius;
  double H = sqrt(top * top - b * b);
  double rho = H * unit_from_coord(uvwz[3], a->sz.r);
  double r = sqrt(rho * rho + b * b);
  double mu;
  int hit;
  if (uvwz[2] < 0.5) {
    double d_min = r - b, d_max = rho;
    double d = d_min + (d_max - d_min) * unit_from_coord(1.0 - 2.0 * uvwz[2], a->sz.mu / 2);
    mu = d == 0.0 ? -1.0 : clamp_cos(-(rho * rho + d * d) / (2.0 * r * d));
    hit = 1;
  } else {
    double d_min = top - r, d_max = rho + H;
    double d = d_min + (d_max - d_min) * unit_from_coord(2.0 * uvwz[2] - 1.0, a->sz.mu / 2);
    mu = d == 0.0 ? 1.0 : clamp_cos((H * H - rho * rho - d * d) / (2.0 * r * d));
    hit = 0;
  }
  double x_mu_s = unit_from_coord(uvwz[1], a->sz.mu_s);
  double d_min = top - b, d_max = H;
  double D = paso_distance_to_top(a, b, a->mu_s_min);
  double A = (D - d_min) / (d_max - d_min);
  double aa = (A - x_mu_s * A) / (1.0 + x_mu_s * A);
  double d = d_min + (aa < A ? aa : A) * (d_max - d_min);
  double mu_s = d == 0.0 ? 1.0 : clamp_cos((H * H - d * d) / (2.0 * b * d));
  out[0] = r;
  out[1] = mu;
  out[2] = mu_s;
  out[3] = clamp_cos(uvwz[0] * 2.0 - 1.0);
  out[4] = hit;
}
void paso_rmumusnu_from_frag_coord(const Atm* a, double x, double y, double z, double* out) {
  /* f.glsl:905-926: the packed x axis holds nu (slow) and mu_s (fast) */
  double fnu = floor(x / a->sz.mu_s);
  double fmus = x - a->sz.mu_s * floor(x / a->sz.mu_s); /* mod(), scalar.h:285-288 */
  double uvwz[4] = {fnu / (a->sz.nu - 1), fmus / a->sz.mu_s, y / a->sz.mu, z / a->sz.r};
  paso_rmumusnu_from_scattering_uvwz(a, uvwz, out);
  double mu = out[1], mu_s = out[2];
  double s = sqrt((1.0 - mu * mu) * (1.0 - mu_s * mu_s));
  out[3] = clampd(out[3], mu * mu_s - s, mu * mu_s + s);
}

/* ---- 4-D lookup = nu-lerp of two trilinear fetches (f.glsl:958-976) ------------------------ */
void paso_get_scattering(const Atm* a, const double* tab, double r, double mu, double mu_s,
                         double nu, int hit, double* out) {
  double uvwz[4];
  paso_scattering_uvwz_from_rmumusnu(a, r, mu, mu_s, nu, hit, uvwz);
  double tcx = uvwz[0] * (a->sz.nu - 1);
  double tx = floor(tcx);
  double f = tcx - tx;
  int nx = a->sz.nu * a->sz.mu_s;
  fetch3(tab, nx, a->sz.mu, a->sz.r, a->nc, (tx + uvwz[1]) / a->sz.nu, uvwz[2], uvwz[3], 1.0 - f, 0, out);
  fetch3(tab, nx, a->sz.mu, a->sz.r, a->nc, (tx + 1.0 + uvwz[1]) / a->sz.nu, uvwz[2], uvwz[3], f, 1, out);
}
/* order-dispatching variant (f.glsl:987-1009) */
static void scattering_of_order(const Atm* a, const double* dR, const double* dM, const double* dS,
                                double r, double mu, double mu_s, double nu, int hit, int order,
                                double* out) {
  if (order == 1) {
    double ray[PASO_MAX_CHANNELS], mie[PASO_MAX_CHANNELS];
    paso_get_scattering(a, dR, r, mu, mu_s, nu, hit, ray);
    paso_get_scattering(a, dM, r, mu, mu_s, nu, hit, mie);
    double pr = paso_rayleigh_phase(nu), pm = paso_mie_phase(a->mie_g, nu);
    for (int c = 0; c < a->nc; ++c) out[c] = ray[c] * pr + mie[c] * pm;
  } else {
    paso_get_scattering(a, dS, r, mu, mu_s, nu, hit, out);
  }
}

/* ---- irradiance table (f.glsl:1524-1601) --------------------------------------------------- */
void paso_irradiance_uv_from_rmus(const Atm* a, double r, double mu_s, double* uv) {
  double x_r = (r - a->bottom_radius) / (a->top_radius - a->bottom_radius);
  uv[0] = coord_from_unit(mu_s * 0.5 + 0.5, a->sz.e_w);
  uv[1] = coord_from_unit(x_r, a->sz.e_h);
}
void paso_rmus_from_irradiance_uv(const Atm* a, double u, double v, double* rmus) {
  rmus[0] = a->bottom_radius + unit_from_coord(v, a->sz.e_h) * (a->top_radius - a->bottom_radius);
  rmus[1] = clamp_cos(2.0 * unit_from_coord(u, a->sz.e_w) - 1.0);
}
void paso_get_irradiance(const Atm* a, const double* E, double r, double mu_s, double* out) {
  double uv[2];
  paso_irradiance_uv_from_rmus(a, r, mu_s, uv);
  fetch2(E, a->sz.e_w, a->sz.e_h, a->nc, uv[0], uv[1], out);
}

/* ---- scattering density: sphere integral, 16 theta x 32 phi (f.glsl:1163-1260) ------------- */
void paso_scattering_density_point(const Atm* a, const double* T, const double* dR,
                                   const double* dM, const double* dS, const double* dE, double r,
                                   double mu, double mu_s, double nu, int order, double* out) {
  const int n = 16;
  const double dphi = PI_D / n, dtheta = PI_D / n;
  double wx = sqrt(1.0 - mu * mu), wz = mu;
  double sx = wx == 0.0 ? 0.0 : (nu - mu * mu_s) / wx;
  double sy = sqrt(pos(1.0 - sx * sx - mu_s * mu_s));
  double sz = mu_s;
  double dens_r = paso_profile_density(a, 0, r - a->bottom_radius);
  double dens_m = paso_profile_density(a, 1, r - a->bottom_radius);
  double t_ground[PASO_MAX_CHANNELS], albedo[PASO_MAX_CHANNELS];
  double incident[PASO_MAX_CHANNELS], e_ground[PASO_MAX_CHANNELS];
  for (int c = 0; c < a->nc; ++c) out[c] = 0.0;
  for (int l = 0; l < n; ++l) {
    double theta = (l + 0.5) * dtheta;
    double ct = cos(theta), st = sin(theta);
    int hit = paso_ray_intersects_ground(a, r, ct);
    double d_ground = 0.0;
    for (int c = 0; c < a->nc; ++c) t_ground[c] = albedo[c] = 0.0;
    if (hit) {
      d_ground = paso_distance_to_bottom(a, r, ct);
      paso_get_transmittance(a, T, r, ct, d_ground, 1, t_ground);
      for (int c = 0; c < a->nc; ++c) albedo[c] = a->ground_albedo[c];
    }
    for (int m = 0; m < 2 * n; ++m) {
      double phi = (m + 0.5) * dphi;
      double ix = cos(phi) * st, iy = sin(phi) * st, iz = ct;
      double domega = dtheta * dphi * sin(theta);
      double nu1 = sx * ix + sy * iy + sz * iz;
      scattering_of_order(a, dR, dM, dS, r, iz, mu_s, nu1, hit, order - 1, incident);
      /* ground bounce: normal of the ground point hit by the ray */
      double gx = ix * d_ground, gy = iy * d_ground, gz = r + iz * d_ground;
      double gl = sqrt(gx * gx + gy * gy + gz * gz);
      double cos_g = (gx * sx + gy * sy + gz * sz) / gl;
      paso_get_irradiance(a, dE, a->bottom_radius, cos_g, e_ground);
      double nu2 = wx * ix + wz * iz;
      double pr = paso_rayleigh_phase(nu2), pm = paso_mie_phase(a->mie_g, nu2);
      for (int c = 0; c < a->nc; ++c) {
        double li = incident[c] + t_ground[c] * albedo[c] * (1.0 / PI_D) * e_ground[c];
        out[c] += li * (a->rayleigh_scattering[c] * dens_r * pr + a->mie_scattering[c] * dens_m * pm) *
                  domega;
      }
    }
  }
}

/* ---- multiple scattering: 50-interval ray march (f.glsl:1285-1330) ------------------------- */
void paso_multiple_scattering_point(const Atm* a, const double* T, const double* dJ, double r,
                                    double mu, double mu_s, double nu, int hit, double* out) {
  const int n = 50;
  double dx = distance_to_nearest_boundary(a, r, mu, hit) / n;
  double j_i[PASO_MAX_CHANNELS], t_i[PASO_MAX_CHANNELS];
  for (int c = 0; c < a->nc; ++c) out[c] = 0.0;
  for (int i = 0; i <= n; ++i) {
    double d = i * dx;
    double r_i = clamp_radius(a, sqrt(d * d + 2.0 * r * mu * d + r * r));
    double mu_i = clamp_cos((r * mu + d) / r_i);
    double mu_s_i = clamp_cos((r * mu_s + d * nu) / r_i);
    paso_get_scattering(a, dJ, r_i, mu_i, mu_s_i, nu, hit, j_i);
    paso_get_transmittance(a, T, r, mu, d, hit, t_i);
    double w = (i == 0 || i == n) ? 0.5 : 1.0;
    for (int c = 0; c < a->nc; ++c) out[c] += j_i[c] * t_i[c] * dx * w;
  }
}

/* ---- irradiance integrals (f.glsl:1443-1511) ----------------------------------------------- */
void paso_direct_irradiance_point(const Atm* a, const double* T, double r, double mu_s,
                                  double* out) {
  double alpha = a->sun_angular_radius;
  double f = mu_s < -alpha ? 0.0
                           : (mu_s > alpha ? mu_s : (mu_s + alpha) * (mu_s + alpha) / (4.0 * alpha));
  transmittance_to_top(a, T, r, mu_s, out);
  for (int c = 0; c < a->nc; ++c) out[c] = a->solar_irradiance[c] * out[c] * f;
}
void paso_indirect_irradiance_point(const Atm* a, const double* dR, const double* dM,
                                    const double* dS, double r, double mu_s, int order,
                                    double* out) {
  const int n = 32;
  const double dphi = PI_D / n, dtheta = PI_D / n;
  double sx = sqrt(1.0 - mu_s * mu_s), sz = mu_s;
  double rad[PASO_MAX_CHANNELS];
  for (int c = 0; c < a->nc; ++c) out[c] = 0.0;
  for (int j = 0; j < n / 2; ++j) {
    double theta = (j + 0.5) * dtheta;
    for (int i = 0; i < 2 * n; ++i) {
      double phi = (i + 0.5) * dphi;
      double wx = cos(phi) * sin(theta), wz = cos(theta);
      double domega = dtheta * dphi * sin(theta);
      double nu = wx * sx + wz * sz;
      scattering_of_order(a, dR, dM, dS, r, wz, mu_s, nu, 0, order, rad);
      for (int c = 0; c < a->nc; ++c) out[c] += rad[c] * wz * domega;
    }
  }
}


/* ---- row-parallel runner (plain pthreads; PASO_THREADS overrides the core count) ----------- */
typedef void (*row_fn)(void* ctx, int row);
typedef struct { row_fn fn; void* ctx; atomic_int next; int end; } RowPool;
static void* row_worker(void* arg) {
  RowPool* p = (RowPool*)arg;
  for (;;) {
    int row = atomic_fetch_add(&p->next, 1);
    if (row >= p->end) return NULL;
    p->fn(p->ctx, row);
  }
}
static int paso_thread_count(void) {
  const char* e = getenv("PASO_THREADS");
  int n = e ? atoi(e) : (int)sysconf(_SC_NPROCESSORS_ONLN);
  return n < 1 ? 1 : (n > 256 ? 256 : n);
}
static void for_rows(int begin, int end, row_fn fn, void* ctx) {
  RowPool pool = {fn, ctx, begin, end};
  int n = paso_thread_count();
  if (end - begin < n) n = end - begin;
  if (n <= 1) { row_worker(&pool); return; }
  pthread_t th[256];
  int started = 0;
  for (int t = 0; t < n; ++t) started += pthread_create(&th[started], NULL, row_worker, &pool) == 0;
  if (started == 0) row_worker(&pool);
  for (int t = 0; t < started; ++t) pthread_join(th[t], NULL);
}

/* ---- whole-table passes (reference/model.cc:140-237; texel centres at +0.5) ---------------- */
static size_t texels2(int w, int h) { return (size_t)w * h; }
static size_t texels3(const Atm* a) { return (size_t)a->sz.nu * a->sz.mu_s * a->sz.mu * a->sz.r; }
static int bad(const Atm* a) { return a == NULL || a->nc < 1 || a->nc > PASO_MAX_CHANNELS; }

typedef struct {
  const Atm* a;
  const double *T, *dR, *dM, *dS, *dE, *dJ;
  double *out0, *out1, *nu_out;
  int order;
} PassCtx;

static void row_transmittance(void* vctx, int j) {
  PassCtx* x = (PassCtx*)vctx;
  const Atm* a = x->a;
  int w = a->sz.t_w, h = a->sz.t_h;
  size_t plane = texels2(w, h);
  double v[PASO_MAX_CHANNELS], rmu[2];
  for (int i = 0; i < w; ++i) {
    paso_rmu_from_transmittance_uv(a, (i + 0.5) / w, (j + 0.5) / h, rmu); /* f.glsl:454-463 */
    paso_compute_transmittance_to_top(a, rmu[0], rmu[1], v);
    for (int c = 0; c < a->nc; ++c) x->out0[c * plane + i + (size_t)w * j] = v[c];
  }
}
int paso_transmittance(const Atm* a, double* T, int row_begin, int row_end) {
  if (bad(a)) return 1;
  PassCtx x = {0};
  x.a = a; x.out0 = T;
  for_rows(row_begin, row_end, row_transmittance, &x);
  return 0;
}

static void row_direct_irradiance(void* vctx, int j) {
  PassCtx* x = (PassCtx*)vctx;
  const Atm* a = x->a;
  int w = a->sz.e_w, h = a->sz.e_h;
  size_t plane = texels2(w, h);
  double v[PASO_MAX_CHANNELS], rmus[2];
  for (int i = 0; i < w; ++i) {
    paso_rmus_from_irradiance_uv(a, (i + 0.5) / w, (j + 0.5) / h, rmus); /* f.glsl:1558-1567 */
    paso_direct_irradiance_point(a, x->T, rmus[0], rmus[1], v);
    for (int c = 0; c < a->nc; ++c) x->out0[c * plane + i + (size_t)w * j] = v[c];
  }
}
int paso_direct_irradiance(const Atm* a, const double* T, double* dE, int row_begin, int row_end) {
  if (bad(a)) return 1;
  PassCtx x = {0};
  x.a = a; x.T = T; x.out0 = dE;
  for_rows(row_begin, row_end, row_direct_irradiance, &x);
  return 0;
}

/* parallel over texels here (only e_h rows exist): "row" = j * e_w + i */
static void texel_indirect_irradiance(void* vctx, int t) {
  PassCtx* x = (PassCtx*)vctx;
  const Atm* a = x->a;
  int w = a->sz.e_w, h = a->sz.e_h;
  size_t plane = texels2(w, h);
  int j = t / w, i = t % w;
  double v[PASO_MAX_CHANNELS], rmus[2];
  paso_rmus_from_irradiance_uv(a, (i + 0.5) / w, (j + 0.5) / h, rmus); /* f.glsl:1573-1586 */
  paso_indirect_irradiance_point(a, x->dR, x->dM, x->dS, rmus[0], rmus[1], x->order, v);
  for (int c = 0; c < a->nc; ++c) x->out0[c * plane + t] = v[c];
}
int paso_indirect_irradiance(const Atm* a, const double* dR, const double* dM, const double* dS,
                             int order, double* dE, int row_begin, int row_end) {
  if (bad(a) || order < 1) return 1;
  PassCtx x = {0};
  x.a = a; x.dR = dR; x.dM = dM; x.dS = dS; x.order = order; x.out0 = dE;
  for_rows(row_begin * a->sz.e_w, row_end * a->sz.e_w, texel_indirect_irradiance, &x);
  return 0;
}

static void row_single_scattering(void* vctx, int row) {
  PassCtx* x = (PassCtx*)vctx;
  const Atm* a = x->a;
  int w = a->sz.nu * a->sz.mu_s, h = a->sz.mu;
  size_t plane = texels3(a);
  int k = row / h, j = row % h;
  double p[5], ray[PASO_MAX_CHANNELS], mie[PASO_MAX_CHANNELS];
  for (int i = 0; i < w; ++i) {
    paso_rmumusnu_from_frag_coord(a, i + 0.5, j + 0.5, k + 0.5, p); /* f.glsl:933-945 */
    paso_single_scattering_point(a, x->T, p[0], p[1], p[2], p[3], (int)p[4], ray, mie);
    size_t t = i + (size_t)w * (j + (size_t)h * k);
    for (int c = 0; c < a->nc; ++c) {
      x->out0[c * plane + t] = ray[c];
      x->out1[c * plane + t] = mie[c];
    }
  }
}
int paso_single_scattering(const Atm* a, const double* T, double* dR, double* dM, int row_begin,
                           int row_end) {
  if (bad(a)) return 1;
  PassCtx x = {0};
  x.a = a; x.T = T; x.out0 = dR; x.out1 = dM;
  for_rows(row_begin, row_end, row_single_scattering, &x);
  return 0;
}

static void row_scattering_density(void* vctx, int row) {
  PassCtx* x = (PassCtx*)vctx;
  const Atm* a = x->a;
  int w = a->sz.nu * a->sz.mu_s, h = a->sz.mu;
  size_t plane = texels3(a);
  int k = row / h, j = row % h;
  double p[5], v[PASO_MAX_CHANNELS];
  for (int i = 0; i < w; ++i) {
    paso_rmumusnu_from_frag_coord(a, i + 0.5, j + 0.5, k + 0.5, p); /* f.glsl:1348-1367 */
    paso_scattering_density_point(a, x->T, x->dR, x->dM, x->dS, x->dE, p[0], p[1], p[2], p[3],
                                  x->order, v);
    size_t t = i + (size_t)w * (j + (size_t)h * k);
    for (int c = 0; c < a->nc; ++c) x->out0[c * plane + t] = v[c];
  }
}
int paso_scattering_density(const Atm* a, const double* T, const double* dR, const double* dM,
                            const double* dS, const double* dE, int order, double* dJ,
                            int row_begin, int row_end) {
  if (bad(a) || order < 2) return 1;
  PassCtx x = {0};
  x.a = a; x.T = T; x.dR = dR; x.dM = dM; x.dS = dS; x.dE = dE; x.order = order; x.out0 = dJ;
  for_rows(row_begin, row_end, row_scattering_density, &x);
  return 0;
}

static void row_multiple_scattering(void* vctx, int row) {
  PassCtx* x = (PassCtx*)vctx;
  const Atm* a = x->a;
  int w = a->sz.nu * a->sz.mu_s, h = a->sz.mu;
  size_t plane = texels3(a);
  int k = row / h, j = row % h;
  double p[5], v[PASO_MAX_CHANNELS];
  for (int i = 0; i < w; ++i) {
    paso_rmumusnu_from_frag_coord(a, i + 0.5, j + 0.5, k + 0.5, p); /* f.glsl:1369-1383 */
    paso_multiple_scattering_point(a, x->T, x->dJ, p[0], p[1], p[2], p[3], (int)p[4], v);
    size_t t = i + (size_t)w * (j + (size_t)h * k);
    for (int c = 0; c < a->nc; ++c) x->out0[c * plane + t] = v[c];
    if (x->nu_out) x->nu_out[t] = p[3];
  }
}
int paso_multiple_scattering(const Atm* a, const double* T, const double* dJ, double* dS,
                             double* nu_out, int row_begin, int row_end) {
  if (bad(a)) return 1;
  PassCtx x = {0};
  x.a = a; x.T = T; x.dJ = dJ; x.out0 = dS; x.nu_out = nu_out;
  for_rows(row_begin, row_end, row_multiple_scattering, &x);
  return 0;
}

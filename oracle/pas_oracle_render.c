/* TEST INFRASTRUCTURE ONLY -- the parity oracle for the render-time lookups (SURVEY.md section 8f,
 * rank 1). See pas_oracle.h for the rules: never linked into or called from the product path.
 *
 * Plain-C, double-precision restatement of
 *   GetExtrapolatedSingleMieScattering   atmosphere/functions.glsl:1634-1646
 *   GetCombinedScattering                atmosphere/functions.glsl:1658-1690
 *   GetSkyRadiance                       atmosphere/functions.glsl:1705-1769
 *   GetSkyRadianceToPoint                atmosphere/functions.glsl:1787-1863
 *   GetSunAndSkyIrradiance               atmosphere/functions.glsl:1878-1896
 *   the luminance wrappers               atmosphere/model.cc:221-281 (kAtmosphereShader)
 *   the test scene GetViewRayRadiance    atmosphere/reference/model_test.glsl:66-348
 *   the view rays and the tone map       atmosphere/reference/model_test.cc:436-477, 688-731
 * for any channel count. The separate-Mie path is pinned against the UNMODIFIED reference
 * (oracle/_ref, pasref_render_scene) by tests/test_oracle_render.py; the combined-texture path
 * exists in GLSL only (functions.glsl:1625-1646) and is restated from the text.
 */
#include <math.h>
#include <stddef.h>
#include <string.h>

#include "pas_oracle.h"

typedef paso_atmosphere Atm;
typedef paso_render_tables Tab;
#define PI_D 3.14159265358979323846

static inline double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static inline double clampd(double x, double lo, double hi) { return x < lo ? lo : (x > hi ? hi : x); }
static inline double smoothstep_d(double e0, double e1, double x) {
  x = clampd((x - e0) / (e1 - e0), 0.0, 1.0);
  return x * x * (3.0 - 2.0 * x);
}

/* functions.glsl:1634-1646; `scattering` holds the nc Rayleigh+multiple channels, mie_red the
 * alpha channel. Channel 0 is "r". */
static void extrapolate_single_mie(const Atm* a, const double* scattering, double mie_red, double* out) {
  if (scattering[0] <= 0.0) {
    for (int c = 0; c < a->nc; ++c) out[c] = 0.0;
    return;
  }
  for (int c = 0; c < a->nc; ++c) {
    out[c] = scattering[c] * mie_red / scattering[0] *
             (a->rayleigh_scattering[0] / a->mie_scattering[0]) *
             (a->mie_scattering[c] / a->rayleigh_scattering[c]);
  }
}

/* functions.glsl:1658-1690. In combined mode *mie_red receives the interpolated alpha channel and
 * single_mie the extrapolated spectrum. */
static void combined_scattering(const Atm* a, const Tab* t, double r, double mu, double mu_s, double nu,
                                int hit, double* scattering, double* single_mie) {
  paso_get_scattering(a, t->scattering, r, mu, mu_s, nu, hit, scattering);
  if (t->single_mie != NULL) {
    paso_get_scattering(a, t->single_mie, r, mu, mu_s, nu, hit, single_mie);
  } else {
    Atm one = *a;
    one.nc = 1;
    double alpha;
    paso_get_scattering(&one, t->scattering_alpha, r, mu, mu_s, nu, hit, &alpha);
    extrapolate_single_mie(a, scattering, alpha, single_mie);
  }
}

static int sky_radiance_radiometric(const Atm* a, const Tab* t, const double* camera_in,
                                    const double* view_ray, double shadow_length,
                                    const double* sun_direction, double* radiance, double* transmittance) {
  double camera[3] = {camera_in[0], camera_in[1], camera_in[2]};
  double r = sqrt(dot3(camera, camera));
  double rmu = dot3(camera, view_ray);
  double dist_top = -rmu - sqrt(rmu * rmu - r * r + a->top_radius * a->top_radius);
  if (dist_top > 0.0) {
    for (int i = 0; i < 3; ++i) camera[i] += view_ray[i] * dist_top;
    r = a->top_radius;
    rmu += dist_top;
  } else if (r > a->top_radius) {
    for (int c = 0; c < a->nc; ++c) { transmittance[c] = 1.0; radiance[c] = 0.0; }
    return 0;
  }
  double mu = rmu / r;
  double mu_s = dot3(camera, sun_direction) / r;
  double nu = dot3(view_ray, sun_direction);
  int hit = paso_ray_intersects_ground(a, r, mu);
  if (hit) {
    for (int c = 0; c < a->nc; ++c) transmittance[c] = 0.0;
  } else {
    paso_get_transmittance_to_top(a, t->transmittance, r, mu, transmittance);
  }
  double scattering[PASO_MAX_CHANNELS], single_mie[PASO_MAX_CHANNELS];
  if (shadow_length == 0.0) {
    combined_scattering(a, t, r, mu, mu_s, nu, hit, scattering, single_mie);
  } else {
    double d = shadow_length;
    double r_p = clampd(sqrt(d * d + 2.0 * r * mu * d + r * r), a->bottom_radius, a->top_radius);
    double mu_p = (r * mu + d) / r_p;
    double mu_s_p = (r * mu_s + d * nu) / r_p;
    combined_scattering(a, t, r_p, mu_p, mu_s_p, nu, hit, scattering, single_mie);
    double shadow_t[PASO_MAX_CHANNELS];
    paso_get_transmittance(a, t->transmittance, r, mu, shadow_length, hit, shadow_t);
    for (int c = 0; c < a->nc; ++c) {
      scattering[c] *= shadow_t[c];
      single_mie[c] *= shadow_t[c];
    }
  }
  double pr = paso_rayleigh_phase(nu), pm = paso_mie_phase(a->mie_g, nu);
  for (int c = 0; c < a->nc; ++c) radiance[c] = scattering[c] * pr + single_mie[c] * pm;
  return 0;
}

static int sky_radiance_to_point_radiometric(const Atm* a, const Tab* t, const double* camera_in,
                                             const double* point, double shadow_length,
                                             const double* sun_direction, double* radiance,
                                             double* transmittance) {
  double camera[3] = {camera_in[0], camera_in[1], camera_in[2]};
  double view_ray[3] = {point[0] - camera[0], point[1] - camera[1], point[2] - camera[2]};
  double len = sqrt(dot3(view_ray, view_ray));
  for (int i = 0; i < 3; ++i) view_ray[i] /= len;
  double r = sqrt(dot3(camera, camera));
  double rmu = dot3(camera, view_ray);
  double dist_top = -rmu - sqrt(rmu * rmu - r * r + a->top_radius * a->top_radius);
  if (dist_top > 0.0) {
    for (int i = 0; i < 3; ++i) camera[i] += view_ray[i] * dist_top;
    r = a->top_radius;
    rmu += dist_top;
  }
  double mu = rmu / r;
  double mu_s = dot3(camera, sun_direction) / r;
  double nu = dot3(view_ray, sun_direction);
  double pc[3] = {point[0] - camera[0], point[1] - camera[1], point[2] - camera[2]};
  double d = sqrt(dot3(pc, pc));
  int hit = paso_ray_intersects_ground(a, r, mu);
  paso_get_transmittance(a, t->transmittance, r, mu, d, hit, transmittance);
  double scattering[PASO_MAX_CHANNELS], single_mie[PASO_MAX_CHANNELS];
  combined_scattering(a, t, r, mu, mu_s, nu, hit, scattering, single_mie);
  d = fmax(d - shadow_length, 0.0);
  double r_p = clampd(sqrt(d * d + 2.0 * r * mu * d + r * r), a->bottom_radius, a->top_radius);
  double mu_p = (r * mu + d) / r_p;
  double mu_s_p = (r * mu_s + d * nu) / r_p;
  double scattering_p[PASO_MAX_CHANNELS], single_mie_p[PASO_MAX_CHANNELS];
  combined_scattering(a, t, r_p, mu_p, mu_s_p, nu, hit, scattering_p, single_mie_p);
  double shadow_t[PASO_MAX_CHANNELS];
  memcpy(shadow_t, transmittance, sizeof(double) * a->nc);
  if (shadow_length > 0.0) paso_get_transmittance(a, t->transmittance, r, mu, d, hit, shadow_t);
  for (int c = 0; c < a->nc; ++c) {
    scattering[c] -= shadow_t[c] * scattering_p[c];
    single_mie[c] -= shadow_t[c] * single_mie_p[c];
  }
  if (t->single_mie == NULL) {
    /* functions.glsl:1851-1854: re-extrapolate from the differenced red channel. In combined mode
     * single_mie[0] before this line is the difference of the two extrapolated reds, which is the
     * differenced alpha only up to the extrapolation's non-linearity; the GLSL uses
     * single_mie_scattering.r as it stands, so do we. */
    double mie_red = single_mie[0];
    extrapolate_single_mie(a, scattering, mie_red, single_mie);
  }
  double fade = smoothstep_d(0.0, 0.01, mu_s);
  double pr = paso_rayleigh_phase(nu), pm = paso_mie_phase(a->mie_g, nu);
  for (int c = 0; c < a->nc; ++c) radiance[c] = scattering[c] * pr + single_mie[c] * fade * pm;
  return 0;
}

static void sun_and_sky_irradiance_radiometric(const Atm* a, const Tab* t, const double* point,
                                               const double* normal, const double* sun_direction,
                                               double* sun_irradiance, double* sky_irradiance) {
  double r = sqrt(dot3(point, point));
  double mu_s = dot3(point, sun_direction) / r;
  paso_get_irradiance(a, t->irradiance, r, mu_s, sky_irradiance);
  double k = (1.0 + dot3(normal, point) / r) * 0.5;
  double ts[PASO_MAX_CHANNELS];
  paso_get_transmittance_to_sun(a, t->transmittance, r, mu_s, ts);
  double cosine = fmax(dot3(normal, sun_direction), 0.0);
  for (int c = 0; c < a->nc; ++c) {
    sky_irradiance[c] *= k;
    sun_irradiance[c] = a->solar_irradiance[c] * ts[c] * cosine;
  }
}

/* ---- public entry points: radiance or luminance according to t->sky_k / t->sun_k ------------ */
int paso_sky_radiance(const Atm* a, const Tab* t, const double* camera, const double* view_ray,
                      double shadow_length, const double* sun_direction, double* radiance,
                      double* transmittance) {
  sky_radiance_radiometric(a, t, camera, view_ray, shadow_length, sun_direction, radiance, transmittance);
  for (int c = 0; c < a->nc; ++c) radiance[c] *= t->sky_k[c];
  return 0;
}
int paso_sky_radiance_to_point(const Atm* a, const Tab* t, const double* camera, const double* point,
                               double shadow_length, const double* sun_direction, double* radiance,
                               double* transmittance) {
  sky_radiance_to_point_radiometric(a, t, camera, point, shadow_length, sun_direction, radiance, transmittance);
  for (int c = 0; c < a->nc; ++c) radiance[c] *= t->sky_k[c];
  return 0;
}
int paso_sun_and_sky_irradiance(const Atm* a, const Tab* t, const double* point, const double* normal,
                                const double* sun_direction, double* sun_irradiance,
                                double* sky_irradiance) {
  sun_and_sky_irradiance_radiometric(a, t, point, normal, sun_direction, sun_irradiance, sky_irradiance);
  for (int c = 0; c < a->nc; ++c) {
    sky_irradiance[c] *= t->sky_k[c];
    sun_irradiance[c] *= t->sun_k[c];
  }
  return 0;
}
void paso_solar_radiance(const Atm* a, const Tab* t, double* out) {
  /* GL model: E / (pi alpha^2) (model.cc:228-231); CPU model: E / (2 pi (1 - cos alpha))
   * (reference/model.cc:255-259) */
  double omega = t->gl_solar_radiance ? PI_D * a->sun_angular_radius * a->sun_angular_radius
                                      : 2.0 * PI_D * (1.0 - cos(a->sun_angular_radius));
  for (int c = 0; c < a->nc; ++c) out[c] = a->solar_irradiance[c] / omega * t->sun_k[c];
}

/* ---- the test scene (model_test.glsl) ------------------------------------------------------- */
static double sun_visibility(const paso_scene* s, const double* point, const double* sun_direction) {
  double p[3] = {point[0] - s->sphere_center[0], point[1] - s->sphere_center[1], point[2] - s->sphere_center[2]};
  double p_dot_v = dot3(p, sun_direction), p_dot_p = dot3(p, p);
  double d2 = p_dot_p - p_dot_v * p_dot_v;
  double dist = -p_dot_v - sqrt(s->sphere_radius * s->sphere_radius - d2);
  if (dist > 0.0) {
    double ray_sphere_distance = s->sphere_radius - sqrt(d2);
    double ang = -ray_sphere_distance / p_dot_v;
    return smoothstep_d(1.0, 0.0, ang / s->sun_size[0]);
  }
  return 1.0;
}
static double sky_visibility(const paso_scene* s, const double* point) {
  double p[3] = {point[0] - s->sphere_center[0], point[1] - s->sphere_center[1], point[2] - s->sphere_center[2]};
  double p_dot_p = dot3(p, p);
  return 1.0 + p[2] / sqrt(p_dot_p) * s->sphere_radius * s->sphere_radius / p_dot_p;
}
static void sphere_shadow_in_out(const paso_scene* s, const double* view_direction, double* d_in, double* d_out) {
  double pos[3] = {s->camera[0] - s->sphere_center[0], s->camera[1] - s->sphere_center[1],
                   s->camera[2] - s->sphere_center[2]};
  double pos_dot_sun = dot3(pos, s->sun_direction);
  double view_dot_sun = dot3(view_direction, s->sun_direction);
  double k = s->sun_size[0], R = s->sphere_radius;
  double l = 1.0 + k * k;
  double a = 1.0 - l * view_dot_sun * view_dot_sun;
  double b = dot3(pos, view_direction) - l * pos_dot_sun * view_dot_sun - k * R * view_dot_sun;
  double c = dot3(pos, pos) - l * pos_dot_sun * pos_dot_sun - 2.0 * k * R * pos_dot_sun - R * R;
  double disc = b * b - a * c;
  if (disc > 0.0) {
    *d_in = fmax(0.0, (-b - sqrt(disc)) / a);
    *d_out = (-b + sqrt(disc)) / a;
    double d_base = -pos_dot_sun / view_dot_sun;
    double d_apex = -(pos_dot_sun + R / k) / view_dot_sun;
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

void paso_view_ray_radiance(const Atm* a, const Tab* t, const paso_scene* s, const double* view_ray,
                            const double* view_ray_diff, double* radiance) {
  const int nc = a->nc;
  double vlen = sqrt(dot3(view_ray, view_ray));
  double v[3] = {view_ray[0] / vlen, view_ray[1] / vlen, view_ray[2] / vlen};
  double fragment_angular_size = sqrt(dot3(view_ray_diff, view_ray_diff)) / vlen;
  double shadow_in, shadow_out;
  sphere_shadow_in_out(s, v, &shadow_in, &shadow_out);
  double cam_e[3] = {s->camera[0] - s->earth_center[0], s->camera[1] - s->earth_center[1],
                     s->camera[2] - s->earth_center[2]};

  /* sphere */
  double p[3] = {s->camera[0] - s->sphere_center[0], s->camera[1] - s->sphere_center[1],
                 s->camera[2] - s->sphere_center[2]};
  double p_dot_v = dot3(p, v), p_dot_p = dot3(p, p);
  double d2 = p_dot_p - p_dot_v * p_dot_v;
  double dist = -p_dot_v - sqrt(s->sphere_radius * s->sphere_radius - d2);
  double sphere_alpha = 0.0;
  double sphere_radiance[PASO_MAX_CHANNELS];
  for (int c = 0; c < nc; ++c) sphere_radiance[c] = 0.0;
  if (dist > 0.0) {
    double ray_sphere_distance = s->sphere_radius - sqrt(d2);
    double ang = -ray_sphere_distance / p_dot_v;
    sphere_alpha = fmin(ang / fragment_angular_size, 1.0);
    double point[3], normal[3], pe[3];
    for (int i = 0; i < 3; ++i) {
      point[i] = s->camera[i] + v[i] * dist;
      normal[i] = point[i] - s->sphere_center[i];
      pe[i] = point[i] - s->earth_center[i];
    }
    double nl = sqrt(dot3(normal, normal));
    for (int i = 0; i < 3; ++i) normal[i] /= nl;
    double sun_e[PASO_MAX_CHANNELS], sky_e[PASO_MAX_CHANNELS];
    paso_sun_and_sky_irradiance(a, t, pe, normal, s->sun_direction, sun_e, sky_e);
    double shadow_length = fmax(0.0, fmin(shadow_out, dist) - shadow_in);
    double tr[PASO_MAX_CHANNELS], in_scatter[PASO_MAX_CHANNELS];
    paso_sky_radiance_to_point(a, t, cam_e, pe, shadow_length, s->sun_direction, in_scatter, tr);
    for (int c = 0; c < nc; ++c) {
      sphere_radiance[c] = s->sphere_albedo[c] * (1.0 / PI_D) * (sun_e[c] + sky_e[c]) * tr[c] + in_scatter[c];
    }
  }

  /* planet */
  p_dot_v = dot3(cam_e, v);
  p_dot_p = dot3(cam_e, cam_e);
  d2 = p_dot_p - p_dot_v * p_dot_v;
  dist = -p_dot_v - sqrt(s->earth_center[2] * s->earth_center[2] - d2);
  double ground_alpha = 0.0;
  double ground_radiance[PASO_MAX_CHANNELS];
  for (int c = 0; c < nc; ++c) ground_radiance[c] = 0.0;
  if (dist > 0.0) {
    double point[3], normal[3], pe[3];
    for (int i = 0; i < 3; ++i) {
      point[i] = s->camera[i] + v[i] * dist;
      pe[i] = point[i] - s->earth_center[i];
      normal[i] = pe[i];
    }
    double nl = sqrt(dot3(normal, normal));
    for (int i = 0; i < 3; ++i) normal[i] /= nl;
    double sun_e[PASO_MAX_CHANNELS], sky_e[PASO_MAX_CHANNELS];
    paso_sun_and_sky_irradiance(a, t, pe, normal, s->sun_direction, sun_e, sky_e);
    double sun_vis = sun_visibility(s, point, s->sun_direction), sky_vis = sky_visibility(s, point);
    double shadow_length = fmax(0.0, fmin(shadow_out, dist) - shadow_in);
    double tr[PASO_MAX_CHANNELS], in_scatter[PASO_MAX_CHANNELS];
    paso_sky_radiance_to_point(a, t, cam_e, pe, shadow_length, s->sun_direction, in_scatter, tr);
    for (int c = 0; c < nc; ++c) {
      ground_radiance[c] =
          s->ground_albedo[c] * (1.0 / PI_D) * (sun_e[c] * sun_vis + sky_e[c] * sky_vis) * tr[c] + in_scatter[c];
    }
    ground_alpha = 1.0;
  }

  /* sky */
  double shadow_length = fmax(0.0, shadow_out - shadow_in);
  double tr[PASO_MAX_CHANNELS];
  paso_sky_radiance(a, t, cam_e, v, shadow_length, s->sun_direction, radiance, tr);
  if (dot3(v, s->sun_direction) > s->sun_size[1]) {
    double solar[PASO_MAX_CHANNELS];
    paso_solar_radiance(a, t, solar);
    for (int c = 0; c < nc; ++c) radiance[c] += tr[c] * solar[c];
  }
  for (int c = 0; c < nc; ++c) {
    radiance[c] = radiance[c] * (1.0 - ground_alpha) + ground_radiance[c] * ground_alpha;
    radiance[c] = radiance[c] * (1.0 - sphere_alpha) + sphere_radiance[c] * sphere_alpha;
  }
}

/* View rays of pixel (i, j), j = 0 at the top (model_test.cc:690-711). */
void paso_pixel_view_ray(const paso_scene* s, int i, int j, double* view_ray, double* view_ray_diff) {
  const double* M = s->model_from_clip;
  double y = 1.0 - 2.0 * (j + 0.5) / s->height, dy = -2.0 / s->height;
  double x = 2.0 * (i + 0.5) / s->width - 1.0, dx = 2.0 / s->width;
  for (int a = 0; a < 3; ++a) {
    view_ray[a] = M[3 * a] * x + M[3 * a + 1] * y + M[3 * a + 2];
    view_ray_diff[a] = M[3 * a] * dx + M[3 * a + 1] * dy;
  }
}

/* out[(j * width + i) * nc + c], rows [row_begin, row_end). */
int paso_render_scene(const Atm* a, const Tab* t, const paso_scene* s, double* out, int row_begin,
                      int row_end) {
  if (a == NULL || t == NULL || s == NULL || out == NULL) return -1;
  for (int j = row_begin; j < row_end; ++j) {
    for (int i = 0; i < s->width; ++i) {
      double view_ray[3], diff[3];
      paso_pixel_view_ray(s, i, j, view_ray, diff);
      paso_view_ray_radiance(a, t, s, view_ray, diff, out + ((size_t)j * s->width + i) * a->nc);
    }
  }
  return 0;
}

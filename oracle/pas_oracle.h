/* TEST INFRASTRUCTURE ONLY -- the parity oracle. Never linked into, imported by or called from
 * the product path (precomputed_atmospheric_scattering_b200/); only tests/, oracle/ scripts,
 * __graft_entry__.smoke() and bench.py's CPU arms use it, and only as the checker.
 *
 * A plain-C, double-precision restatement of the LUT precomputation of
 * ebruneton/precomputed_atmospheric_scattering (atmosphere/functions.glsl:113-1601 as driven by
 * atmosphere/reference/model.cc:140-237), for an arbitrary number of spectral channels and
 * run-time table sizes. It deliberately evaluates every texel the literal way (full 4-D lookups
 * per direction / sample), so it does not share the algebraic shortcuts of the CUDA kernels.
 *
 * Pinned against the reference itself: tests/test_oracle_vs_reference.py and oracle/gen_golden.py
 * compare it with oracle/_ref (the unmodified reference CPU model compiled from /root/reference)
 * and the committed fixtures tests/golden/ were produced by that reference.
 *
 * Table layout everywhere: planar, tab[c * texels + texel], texel = i + nx*(j + ny*k), i.e. the
 * reference's x-fastest order (external/dimensional_types/math/binary_function.h:73-83,
 * ternary_function.h:74-79) with one plane per channel.
 */
#ifndef PAS_ORACLE_H_
#define PAS_ORACLE_H_

#ifdef __cplusplus
extern "C" {
#endif

#define PASO_MAX_CHANNELS 48

typedef struct paso_sizes {
  int t_w, t_h;                 /* transmittance: x = mu, y = r        (constants.h:47-48) */
  int r, mu, mu_s, nu;          /* scattering 4-D sizes               (constants.h:50-53) */
  int e_w, e_h;                 /* irradiance: x = mu_s, y = r         (constants.h:60-61) */
} paso_sizes;

typedef struct paso_atmosphere {
  int nc;                        /* number of spectral channels */
  double solar_irradiance[PASO_MAX_CHANNELS];
  double rayleigh_scattering[PASO_MAX_CHANNELS];
  double mie_scattering[PASO_MAX_CHANNELS];
  double mie_extinction[PASO_MAX_CHANNELS];
  double absorption_extinction[PASO_MAX_CHANNELS];
  double ground_albedo[PASO_MAX_CHANNELS];
  double sun_angular_radius, bottom_radius, top_radius, mie_g, mu_s_min;
  double profiles[3][2][5];      /* rayleigh, mie, absorption; (width, exp_term, exp_scale,
                                    linear_term, constant_term) (definitions.glsl:185-211) */
  paso_sizes sz;
} paso_atmosphere;

/* Whole-table passes. row ranges select texel rows [row_begin, row_end): a row is one j for the
 * 2-D tables, one (k, j) pair (row = k * mu + j) for the 3-D tables. Outputs outside the range are
 * left untouched. All return 0 on success. */
int paso_transmittance(const paso_atmosphere* a, double* T, int row_begin, int row_end);
int paso_direct_irradiance(const paso_atmosphere* a, const double* T, double* dE,
                           int row_begin, int row_end);
int paso_single_scattering(const paso_atmosphere* a, const double* T, double* dR, double* dM,
                           int row_begin, int row_end);
int paso_scattering_density(const paso_atmosphere* a, const double* T, const double* dR,
                            const double* dM, const double* dS, const double* dE, int order,
                            double* dJ, int row_begin, int row_end);
int paso_indirect_irradiance(const paso_atmosphere* a, const double* dR, const double* dM,
                             const double* dS, int order, double* dE, int row_begin, int row_end);
/* nu_out (optional, one value per texel) receives the texel's clamped nu, needed by the
 * 1 / RayleighPhaseFunction(nu) accumulation (reference/model.cc:231-233). */
int paso_multiple_scattering(const paso_atmosphere* a, const double* T, const double* dJ,
                             double* dS, double* nu_out, int row_begin, int row_end);

/* Point functions (for the analytic known-answer tests of reference/functions_test.cc). */
double paso_distance_to_top(const paso_atmosphere* a, double r, double mu);
double paso_distance_to_bottom(const paso_atmosphere* a, double r, double mu);
int paso_ray_intersects_ground(const paso_atmosphere* a, double r, double mu);
double paso_profile_density(const paso_atmosphere* a, int profile, double altitude);
double paso_optical_length_to_top(const paso_atmosphere* a, int profile, double r, double mu);
void paso_compute_transmittance_to_top(const paso_atmosphere* a, double r, double mu, double* out);
double paso_rayleigh_phase(double nu);
double paso_mie_phase(double g, double nu);
void paso_transmittance_uv_from_rmu(const paso_atmosphere* a, double r, double mu, double* uv);
void paso_rmu_from_transmittance_uv(const paso_atmosphere* a, double u, double v, double* rmu);
void paso_scattering_uvwz_from_rmumusnu(const paso_atmosphere* a, double r, double mu, double mu_s,
                                        double nu, int hit, double* uvwz);
void paso_rmumusnu_from_scattering_uvwz(const paso_atmosphere* a, const double* uvwz, double* out5);
void paso_rmumusnu_from_frag_coord(const paso_atmosphere* a, double x, double y, double z,
                                   double* out5);
void paso_irradiance_uv_from_rmus(const paso_atmosphere* a, double r, double mu_s, double* uv);
void paso_rmus_from_irradiance_uv(const paso_atmosphere* a, double u, double v, double* rmus);
void paso_get_transmittance(const paso_atmosphere* a, const double* T, double r, double mu,
                            double d, int hit, double* out);
void paso_get_transmittance_to_sun(const paso_atmosphere* a, const double* T, double r,
                                   double mu_s, double* out);
void paso_get_scattering(const paso_atmosphere* a, const double* tab, double r, double mu,
                         double mu_s, double nu, int hit, double* out);
void paso_get_irradiance(const paso_atmosphere* a, const double* E, double r, double mu_s,
                         double* out);
void paso_single_scattering_point(const paso_atmosphere* a, const double* T, double r, double mu,
                                  double mu_s, double nu, int hit, double* rayleigh, double* mie);
void paso_scattering_density_point(const paso_atmosphere* a, const double* T, const double* dR,
                                   const double* dM, const double* dS, const double* dE, double r,
                                   double mu, double mu_s, double nu, int order, double* out);
void paso_multiple_scattering_point(const paso_atmosphere* a, const double* T, const double* dJ,
                                    double r, double mu, double mu_s, double nu, int hit,
                                    double* out);
void paso_indirect_irradiance_point(const paso_atmosphere* a, const double* dR, const double* dM,
                                    const double* dS, double r, double mu_s, int order,
                                    double* out);
void paso_direct_irradiance_point(const paso_atmosphere* a, const double* T, double r,
                                  double mu_s, double* out);

void paso_get_transmittance_to_top(const paso_atmosphere* a, const double* T, double r, double mu,
                                   double* out);

/* ---- render-time lookups and the test scene (pas_oracle_render.c) --------------------------- */
typedef struct paso_render_tables {
  const double* transmittance;      /* [nc][t_h][t_w] */
  const double* scattering;         /* [nc][r][mu][nu*mu_s]: Rayleigh + multiple scattering */
  const double* single_mie;         /* [nc][...] or NULL = combined textures (functions.glsl:1625) */
  const double* scattering_alpha;   /* [1][...]: red single Mie, used when single_mie == NULL */
  const double* irradiance;         /* [nc][e_h][e_w] */
  double sky_k[PASO_MAX_CHANNELS];  /* SKY_SPECTRAL_RADIANCE_TO_LUMINANCE, or 1 (model.cc:668-686) */
  double sun_k[PASO_MAX_CHANNELS];  /* SUN_SPECTRAL_RADIANCE_TO_LUMINANCE, or 1 */
  int gl_solar_radiance;            /* 1: E/(pi a^2) (model.cc:228); 0: E/(2pi(1-cos a)) (reference/model.cc:255) */
} paso_render_tables;

typedef struct paso_scene {         /* uniforms of reference/model_test.cc:127-134, 436-477 */
  double camera[3], earth_center[3], sun_direction[3];
  double sun_size[2];               /* tan, cos of the sun angular radius */
  double sphere_center[3], sphere_radius;
  double model_from_clip[9];        /* row major */
  double ground_albedo[PASO_MAX_CHANNELS], sphere_albedo[PASO_MAX_CHANNELS];
  int width, height;
} paso_scene;

int paso_sky_radiance(const paso_atmosphere* a, const paso_render_tables* t, const double* camera,
                      const double* view_ray, double shadow_length, const double* sun_direction,
                      double* radiance, double* transmittance);
int paso_sky_radiance_to_point(const paso_atmosphere* a, const paso_render_tables* t,
                               const double* camera, const double* point, double shadow_length,
                               const double* sun_direction, double* radiance, double* transmittance);
int paso_sun_and_sky_irradiance(const paso_atmosphere* a, const paso_render_tables* t,
                                const double* point, const double* normal,
                                const double* sun_direction, double* sun_irradiance,
                                double* sky_irradiance);
void paso_solar_radiance(const paso_atmosphere* a, const paso_render_tables* t, double* out);
void paso_pixel_view_ray(const paso_scene* s, int i, int j, double* view_ray, double* view_ray_diff);
void paso_view_ray_radiance(const paso_atmosphere* a, const paso_render_tables* t,
                            const paso_scene* s, const double* view_ray, const double* view_ray_diff,
                            double* radiance);
/* out[(j * width + i) * nc + c] for rows [row_begin, row_end), j = 0 at the top. */
int paso_render_scene(const paso_atmosphere* a, const paso_render_tables* t, const paso_scene* s,
                      double* out, int row_begin, int row_end);

#ifdef __cplusplus
}
#endif
#endif  /* PAS_ORACLE_H_ */

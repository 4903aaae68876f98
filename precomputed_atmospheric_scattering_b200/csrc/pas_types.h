// Shared host/device parameter blocks of the B200 LUT precompute engine.
//
// Everything wavelength-independent lives in PasGeometry (fp64 masters: the per-texel and
// per-(layer, direction) setup math runs in double so that the cancellation-prone expressions of
// atmosphere/functions.glsl:211-212, 411, 792-793 are exact); everything per spectral channel
// lives in PasSpectrum (at most PAS_MAX_CH channels per launch group).
#ifndef PAS_B200_CSRC_PAS_TYPES_H_
#define PAS_B200_CSRC_PAS_TYPES_H_

#define PAS_MAX_CH 16          // spectral channels processed by one kernel launch
#define PAS_MAX_NU 16          // max SCATTERING_TEXTURE_NU_SIZE supported by the density kernel
#define PAS_DIR_THETA 16       // scattering-density sphere integral: 16 theta x 32 phi
#define PAS_DIR_PHI 32         //   (atmosphere/functions.glsl:1187,1194,1215)
#define PAS_IRR_THETA 16       // indirect-irradiance hemisphere integral: 16 theta x 64 phi
#define PAS_IRR_PHI 64         //   (atmosphere/functions.glsl:1487,1494,1496)
#define PAS_RAY_SAMPLES 50     // ray-march intervals (functions.glsl:707, 1297)
#define PAS_OPTICAL_SAMPLES 500  // optical-length intervals (functions.glsl:281)

// Table layout in HBM. Every multi-channel intermediate table (transmittance, delta Rayleigh / Mie,
// scattering density, delta multiple scattering) is CHANNEL-INTERLEAVED: tab[texel * CP + c] with
// the channel pitch CP = nc rounded up to a multiple of 4 (3 -> 4, 15 -> 16 floats = 64 B per
// texel), texel = x + width * (j + mu_n * k) in the reference's x-fastest order. One texel is then
// one or a few 16-byte vectors, so a table row (all channels) is a single contiguous run: the ray
// march kernels stage rows with LDG.128 / STS.128 and the density kernel reads and writes whole
// texels. The padding channels are kept at zero. The irradiance tables (64 x 16) stay planar.
#define PAS_CHANNEL_PITCH(nc) (((nc) + 3) & ~3)

struct PasSizes {
  int t_w, t_h;              // transmittance table: x = mu, y = r      (constants.h:47-48)
  int r_n, mu_n, mu_s_n, nu_n;  // scattering table 4-D sizes            (constants.h:50-53)
  int e_w, e_h;              // irradiance table: x = mu_s, y = r       (constants.h:60-61)
};

struct PasGeometry {
  PasSizes sz;
  double bottom, top;        // radii, in length units
  double H;                  // sqrt(top^2 - bottom^2)
  double mu_s_min;
  double mus_A;              // "A" of the mu_s mapping (functions.glsl:819-821)
  double sun_angular_radius;
  double mie_g;
  double profiles[3][2][5];  // rayleigh, mie, absorption x 2 layers x
                             // (width, exp_term, exp_scale, linear_term, constant_term)
};

struct PasSpectrum {
  int nc;
  double solar[PAS_MAX_CH];
  double beta_r[PAS_MAX_CH];       // rayleigh scattering
  double beta_m_sca[PAS_MAX_CH];   // mie scattering
  double beta_m_ext[PAS_MAX_CH];   // mie extinction
  double beta_abs[PAS_MAX_CH];     // absorption extinction
  double albedo[PAS_MAX_CH];
  float lum[3][PAS_MAX_CH];        // luminance_from_radiance columns of this group (model.cc:925-943)
};

// Per-(layer k, polar direction l) constants of the scattering-density pass, produced once per
// Init by density_setup_kernel (they depend on the transmittance table only).
struct PasDensityDir {
  float cos_t, sin_t;        // direction (functions.glsl:1195-1197)
  float w_row;               // weight of row j1 in the mu interpolation at (r_k, cos_t)
  int j0, j1;                // bracketing mu rows (clamped)
  int hit;                   // RayIntersectsGround(r_k, cos_t)
  float dg_over_b;           // distance_to_ground / bottom_radius (0 if !hit)
  float pad;
};

#endif  // PAS_B200_CSRC_PAS_TYPES_H_

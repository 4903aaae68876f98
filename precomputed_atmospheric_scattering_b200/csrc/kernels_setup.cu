// Small fp64 passes: transmittance table, direct irradiance, density-pass setup tables, RGBA pack.
// None of them is on the critical path (a few tens of microseconds together); they run in double
// precision so that everything downstream starts from tables that agree with the CPU reference
// to fp32 rounding.
#include "pas_kernels.h"
#include "pas_physics.cuh"

namespace pas {
namespace {

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ComputeTransmittanceToTopAtmosphereBoundaryTexture (functions.glsl:275-320, 427-463).
// One warp per texel; the 501 trapezoid samples are strided over the lanes.
__global__ void __launch_bounds__(256)
transmittance_kernel(const __grid_constant__ PasGeometry g, const __grid_constant__ PasSpectrum s,
                     float* __restrict__ T, const __grid_constant__ PeerTables mirrors, int j_begin,
                     int j_end, const __grid_constant__ RgbExtinction rgb, float* __restrict__ rgba,
                     const __grid_constant__ PeerTables rgba_mirrors) {
  const int texel = j_begin * g.sz.t_w + ((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (texel >= j_end * g.sz.t_w) return;
  const int i = texel % g.sz.t_w, j = texel / g.sz.t_w;
  // texel -> (r, mu), functions.glsl:427-447
  const double x_mu = unit_from_coord((i + 0.5) / g.sz.t_w, g.sz.t_w);
  double r, rho;
  layer_radius(g, (j + 0.5) / g.sz.t_h, g.sz.t_h, &r, &rho);
  const double d_min = g.top - r, d_max = rho + g.H;
  const double d = d_min + x_mu * (d_max - d_min);
  double mu = d == 0.0 ? 1.0 : (g.H * g.H - rho * rho - d * d) / (2.0 * r * d);
  mu = d_clamp(mu, -1.0, 1.0);
  // optical lengths of the three profiles, functions.glsl:275-299
  const double dx = dist_top(g, r, mu) / PAS_OPTICAL_SAMPLES;
  double acc[3] = {0.0, 0.0, 0.0};
  for (int smp = lane; smp <= PAS_OPTICAL_SAMPLES; smp += 32) {
    const double di = smp * dx;
    const double ri = sqrt(di * di + 2.0 * r * mu * di + r * r);
    const double h = ri - g.bottom;
    const double w = (smp == 0 || smp == PAS_OPTICAL_SAMPLES) ? 0.5 : 1.0;
#pragma unroll
    for (int p = 0; p < 3; ++p) acc[p] += profile_density(g.profiles[p], h) * w;
  }
#pragma unroll
  for (int p = 0; p < 3; ++p) acc[p] = warp_sum(acc[p]) * dx;
  const int cp = PAS_CHANNEL_PITCH(s.nc);
  if (lane < cp) {
    float t = 0.f;  // padding channels stay zero
    if (lane < s.nc) {
      const double tau = s.beta_r[lane] * acc[0] + s.beta_m_ext[lane] * acc[1] +
                         s.beta_abs[lane] * acc[2];
      t = (float)exp(-tau);
    }
    T[(size_t)texel * cp + lane] = t;
    // multi-GPU: every rank computes a band of rows and stores it to all the others
    for (int p = 0; p < mirrors.n; ++p) mirrors.tab[p][(size_t)texel * cp + lane] = t;
  }
  // The optical lengths do not depend on the wavelength: the final RGBA transmittance texture at
  // 680 / 550 / 440 nm (model.cc:951-963) costs three more exponentials, on the idle lanes 16..19.
  if (rgba != nullptr && lane >= 16 && lane < 20) {
    const int c = lane - 16;
    float t = 1.0f;  // alpha
    if (c < 3) {
      t = (float)exp(-(rgb.beta_r[c] * acc[0] + rgb.beta_m_ext[c] * acc[1] + rgb.beta_abs[c] * acc[2]));
    }
    rgba[(size_t)texel * 4 + c] = t;
    for (int p = 0; p < rgba_mirrors.n; ++p) rgba_mirrors.tab[p][(size_t)texel * 4 + c] = t;
  }
}

__global__ void pack_rgba_kernel(const float* __restrict__ table, int n, int nc,
                                 float* __restrict__ rgba) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const int cp = PAS_CHANNEL_PITCH(nc);
  float4 v;
  v.x = table[(size_t)t * cp];
  v.y = nc > 1 ? table[(size_t)t * cp + 1] : 0.0f;
  v.z = nc > 2 ? table[(size_t)t * cp + 2] : 0.0f;
  v.w = 1.0f;  // unspecified in the reference (vec3 written to an RGBA target)
  reinterpret_cast<float4*>(rgba)[t] = v;
}

// Bilinear fetch of channel c of the interleaved transmittance table at texel-space (x, y), fp64
// arithmetic on the fp32 table (binary_function.h:103-118).
__device__ __forceinline__ double fetch_t(const float* __restrict__ T, int cp, int c, int w,
                                          const Tap& tx, const Tap& ty) {
  const double a = T[(size_t)(tx.i0 + w * ty.i0) * cp + c], b = T[(size_t)(tx.i1 + w * ty.i0) * cp + c];
  const double e = T[(size_t)(tx.i0 + w * ty.i1) * cp + c], d = T[(size_t)(tx.i1 + w * ty.i1) * cp + c];
  const double wx = tx.w, wy = ty.w;
  return a * ((1.0 - wx) * (1.0 - wy)) + b * (wx * (1.0 - wy)) + e * ((1.0 - wx) * wy) + d * (wx * wy);
}

// ComputeDirectIrradianceTexture (functions.glsl:1443-1461, 1558-1567).
__global__ void direct_irradiance_kernel(const __grid_constant__ PasGeometry g,
                                         const __grid_constant__ PasSpectrum s,
                                         const float* __restrict__ T, float* __restrict__ dE,
                                         FinalTables fin) {
  const int n = g.sz.e_w * g.sz.e_h;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const int i = t % g.sz.e_w, j = t / g.sz.e_w;
  const double r = g.bottom + unit_from_coord((j + 0.5) / g.sz.e_h, g.sz.e_h) * (g.top - g.bottom);
  const double mu_s = d_clamp(2.0 * unit_from_coord((i + 0.5) / g.sz.e_w, g.sz.e_w) - 1.0, -1.0, 1.0);
  const double alpha = g.sun_angular_radius;
  const double f = mu_s < -alpha ? 0.0
                 : (mu_s > alpha ? mu_s : (mu_s + alpha) * (mu_s + alpha) / (4.0 * alpha));
  double x, y;
  transmittance_xy(g, r, mu_s, &x, &y);
  const Tap tx = make_tap(x, g.sz.t_w), ty = make_tap(y, g.sz.t_h);
  const int cp = PAS_CHANNEL_PITCH(s.nc);
  for (int c = 0; c < s.nc; ++c) {
    dE[(size_t)c * n + t] = (float)(s.solar[c] * fetch_t(T, cp, c, g.sz.t_w, tx, ty) * f);
  }
  if (!fin.accumulate && fin.irradiance != nullptr) {
    reinterpret_cast<float4*>(fin.irradiance)[t] = make_float4(0.f, 0.f, 0.f, 1.f);
  }
}

// Per-(layer, direction) tables of the density pass. Thread (k, l).
__global__ void density_setup_kernel(const __grid_constant__ PasGeometry g,
                                     const __grid_constant__ PasSpectrum s,
                                     const float* __restrict__ T, PasDensityDir* __restrict__ dirs,
                                     float* __restrict__ G, float* __restrict__ cR,
                                     float* __restrict__ cM) {
  const int k = blockIdx.x, l = threadIdx.x;
  double r, rho;
  layer_radius(g, (k + 0.5) / g.sz.r_n, g.sz.r_n, &r, &rho);
  if (l < s.nc) {
    const double h = r - g.bottom;
    cR[k * PAS_MAX_CH + l] = (float)(s.beta_r[l] * profile_density(g.profiles[0], h));
    cM[k * PAS_MAX_CH + l] = (float)(s.beta_m_sca[l] * profile_density(g.profiles[1], h));
  }
  if (l >= PAS_DIR_THETA) return;
  const double theta = (l + 0.5) * (kPi / PAS_DIR_THETA);
  const double ct = cos(theta), st = sin(theta);
  const bool hit = hits_ground(g, r, ct);
  PasDensityDir d;
  d.cos_t = (float)ct;
  d.sin_t = (float)st;
  d.hit = hit ? 1 : 0;
  const Tap row = make_tap(scattering_y_from_mu(g, r, rho, ct, hit), g.sz.mu_n);
  d.j0 = row.i0;
  d.j1 = row.i1;
  d.w_row = row.w;
  const double dg = hit ? dist_bottom(g, r, ct) : 0.0;
  d.dg_over_b = (float)(dg / g.bottom);
  d.pad = 0.f;
  dirs[k * PAS_DIR_THETA + l] = d;
  // transmittance to the ground along the ray, GetTransmittance(r, ct, dg, true)
  // (functions.glsl:493-519): T(r_d, -mu_d) / T(r, -mu), capped at 1.
  const int cp = PAS_CHANNEL_PITCH(s.nc);
  float* Gkl = G + (size_t)(k * PAS_DIR_THETA + l) * PAS_MAX_CH;
  if (!hit) {
    for (int c = 0; c < s.nc; ++c) Gkl[c] = 0.f;
    return;
  }
  const double r_d = d_clamp(sqrt(dg * dg + 2.0 * r * ct * dg + r * r), g.bottom, g.top);
  const double mu_d = d_clamp((r * ct + dg) / r_d, -1.0, 1.0);
  double x0, y0, x1, y1;
  transmittance_xy(g, r_d, -mu_d, &x0, &y0);
  transmittance_xy(g, r, -ct, &x1, &y1);
  const Tap ax = make_tap(x0, g.sz.t_w), ay = make_tap(y0, g.sz.t_h);
  const Tap bx = make_tap(x1, g.sz.t_w), by = make_tap(y1, g.sz.t_h);
  for (int c = 0; c < s.nc; ++c) {
    const double t = fmin(fetch_t(T, cp, c, g.sz.t_w, ax, ay) / fetch_t(T, cp, c, g.sz.t_w, bx, by), 1.0);
    Gkl[c] = (float)(t * s.albedo[c] * (1.0 / kPi));
  }
}

// planar [nc][n] <-> interleaved [n][cp] (test hooks and captures present tables planar)
__global__ void interleaved_to_planar_kernel(const float* __restrict__ src, size_t n, int nc, int cp,
                                             float* __restrict__ dst) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  for (int c = 0; c < nc; ++c) dst[(size_t)c * n + t] = src[t * cp + c];
}
__global__ void planar_to_interleaved_kernel(const float* __restrict__ src, size_t n, int nc, int cp,
                                             float* __restrict__ dst) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  for (int c = 0; c < cp; ++c) dst[t * cp + c] = c < nc ? src[(size_t)c * n + t] : 0.f;
}

}  // namespace

cudaError_t launch_interleaved_to_planar(const float* src, size_t n_texels, int nc, float* dst,
                                         cudaStream_t stream) {
  interleaved_to_planar_kernel<<<(unsigned)((n_texels + 255) / 256), 256, 0, stream>>>(
      src, n_texels, nc, PAS_CHANNEL_PITCH(nc), dst);
  return cudaGetLastError();
}

cudaError_t launch_planar_to_interleaved(const float* src, size_t n_texels, int nc, float* dst,
                                         cudaStream_t stream) {
  planar_to_interleaved_kernel<<<(unsigned)((n_texels + 255) / 256), 256, 0, stream>>>(
      src, n_texels, nc, PAS_CHANNEL_PITCH(nc), dst);
  return cudaGetLastError();
}

cudaError_t launch_transmittance(const PasGeometry& g, const PasSpectrum& s, float* T,
                                 cudaStream_t stream, const PasSpectrum* rgb, float* rgba) {
  RgbExtinction e{};
  if (rgb != nullptr && rgba != nullptr) {
    for (int c = 0; c < 3; ++c) {
      e.beta_r[c] = rgb->beta_r[c];
      e.beta_m_ext[c] = rgb->beta_m_ext[c];
      e.beta_abs[c] = rgb->beta_abs[c];
    }
  } else {
    rgba = nullptr;
  }
  const int n = g.sz.t_w * g.sz.t_h;
  const int warps_per_block = 8;
  transmittance_kernel<<<(n + warps_per_block - 1) / warps_per_block, warps_per_block * 32, 0,
                         stream>>>(g, s, T, PeerTables{}, 0, g.sz.t_h, e, rgba, PeerTables{});
  return cudaGetLastError();
}

cudaError_t launch_transmittance_rows(const PasGeometry& g, const PasSpectrum& s, float* T,
                                      const PeerTables& mirrors, int j_begin, int j_end,
                                      cudaStream_t stream, const PasSpectrum* rgb, float* rgba,
                                      const PeerTables* rgba_mirrors) {
  const int n = g.sz.t_w * (j_end - j_begin);
  if (n <= 0) return cudaSuccess;
  RgbExtinction e{};
  if (rgb != nullptr && rgba != nullptr) {
    for (int c = 0; c < 3; ++c) {
      e.beta_r[c] = rgb->beta_r[c];
      e.beta_m_ext[c] = rgb->beta_m_ext[c];
      e.beta_abs[c] = rgb->beta_abs[c];
    }
  } else {
    rgba = nullptr;
  }
  const int warps_per_block = 8;
  transmittance_kernel<<<(n + warps_per_block - 1) / warps_per_block, warps_per_block * 32, 0,
                         stream>>>(g, s, T, mirrors, j_begin, j_end, e, rgba,
                                   rgba != nullptr && rgba_mirrors != nullptr ? *rgba_mirrors : PeerTables{});
  return cudaGetLastError();
}

cudaError_t launch_pack_rgba(const float* planar, int n_texels, int nc, float* rgba,
                             cudaStream_t stream) {
  pack_rgba_kernel<<<(n_texels + 255) / 256, 256, 0, stream>>>(planar, n_texels, nc, rgba);
  return cudaGetLastError();
}

cudaError_t launch_direct_irradiance(const PasGeometry& g, const PasSpectrum& s, const float* T,
                                     float* dE, FinalTables fin, cudaStream_t stream) {
  const int n = g.sz.e_w * g.sz.e_h;
  direct_irradiance_kernel<<<(n + 127) / 128, 128, 0, stream>>>(g, s, T, dE, fin);
  return cudaGetLastError();
}

cudaError_t launch_density_setup(const PasGeometry& g, const PasSpectrum& s, const float* T,
                                 PasDensityDir* dirs, float* G, float* cR, float* cM,
                                 cudaStream_t stream) {
  density_setup_kernel<<<g.sz.r_n, 32, 0, stream>>>(g, s, T, dirs, G, cR, cM);
  return cudaGetLastError();
}

}  // namespace pas

// The two ray-march passes: single scattering (ComputeSingleScatteringTexture,
// atmosphere/functions.glsl:650-730, 933-945) and multiple scattering
// (ComputeMultipleScatteringTexture, functions.glsl:1285-1330, 1369-1383), each with its fused
// luminance / accumulation epilogue (atmosphere/model.cc:142-157, 192-208).
//
// Mapping: one block per (layer k, mu row j); one thread per x = i_nu * mu_s_n + i_mu_s. All
// texels of a block share the ray (r, mu): the 51 sample points, their radii, the transmittance
// along the ray and the (r, mu) interpolation footprint in the source table are identical for
// every thread. So per sample:
//   phase A (once per block, 51 threads, fp64): sample geometry, path transmittance per channel,
//           table taps of the shared axes;
//   stage   (all threads): the shared-axis interpolation is applied ONCE to a whole table row:
//           the 4 (layer, row) corner rows of the channel-interleaved source table (16 KB each at
//           15 channels, contiguous) are read with 128-bit loads from L2, combined in registers
//           and written to shared memory as one row of texels (multiple scattering); the 2
//           transmittance rows bracketing r_i likewise (single scattering);
//   consume (per thread, fp32): only the thread-dependent axes remain (mu_s and nu, resp. the
//           sun-direction mu): 2-4 texel reads (128-bit, XOR-swizzled so that neighbouring
//           texels fall in different banks) instead of 16 L2 gathers per channel.
// The stage buffer is double buffered: one __syncthreads per sample.
#include <cstdlib>

#include "pas_kernels.h"
#include "pas_physics.cuh"

namespace pas {
namespace {

constexpr int kSamples = PAS_RAY_SAMPLES + 1;

// Bilinear fetch of channel c of the interleaved transmittance table, fp64 arithmetic on the fp32
// table (binary_function.h:103-118).
__device__ __forceinline__ double fetch_t(const float* __restrict__ T, int cp, int c, int w,
                                          const Tap& tx, const Tap& ty) {
  const double a = T[(size_t)(tx.i0 + w * ty.i0) * cp + c], b = T[(size_t)(tx.i1 + w * ty.i0) * cp + c];
  const double e = T[(size_t)(tx.i0 + w * ty.i1) * cp + c], d = T[(size_t)(tx.i1 + w * ty.i1) * cp + c];
  const double wx = tx.w, wy = ty.w;
  return a * ((1.0 - wx) * (1.0 - wy)) + b * (wx * (1.0 - wy)) + e * ((1.0 - wx) * wy) + d * (wx * wy);
}

// Shared per-sample record (written in phase A).
struct RaySample {
  float d;          // distance along the ray
  float inv_r;      // 1 / r_i
  // multiple scattering: footprint of (r_i, mu_i) in the source table, as row offsets (in texels)
  // and the four bilinear weights
  int row00, row01, row10, row11;
  float w00, w01, w10, w11;
  // single scattering: sun-lookup geometry at r_i
  float q;          // (top - r_i)(top + r_i)
  float d_min;      // top - r_i
  float x_scale;    // (t_w - 1) / (d_max - d_min)
  float cos_h;      // cosine of the horizon angle at r_i (functions.glsl:556-557)
  float inv_sun_w;  // 1 / (2 sin_h alpha_s)
  float dens_r, dens_m;
  int y0, y1;       // transmittance rows bracketing r_i
  float wy;
};

// Phase A, one thread per sample (fp64). `want_scatter` fills the multiple-scattering fields,
// otherwise the single-scattering ones. Tw[c] = T(r, mu, d_i)[c] * trapezoid weight * dx.
template <int NC>
__device__ void ray_sample_setup(const PasGeometry& g, const float* __restrict__ T, double r,
                                 double rho, double mu, bool hit, double d_end, int i,
                                 bool want_scatter, RaySample* out, float* Tw) {
  constexpr int CP = PAS_CHANNEL_PITCH(NC);
  const double dx = d_end / PAS_RAY_SAMPLES;
  const double d = i * dx;
  const double r_i = d_clamp(sqrt(d * d + 2.0 * r * mu * d + r * r), g.bottom, g.top);
  const double mu_i = d_clamp((r * mu + d) / r_i, -1.0, 1.0);
  const double rho_i = sqrt(d_pos(r_i * r_i - g.bottom * g.bottom));
  RaySample s;
  s.d = (float)d;
  s.inv_r = (float)(1.0 / r_i);
  if (want_scatter) {
    const Tap tk = make_tap(rho_i / g.H * (g.sz.r_n - 1), g.sz.r_n);
    const Tap tj = make_tap(scattering_y_from_mu(g, r_i, rho_i, mu_i, hit), g.sz.mu_n);
    const int width = g.sz.nu_n * g.sz.mu_s_n;
    s.row00 = (tk.i0 * g.sz.mu_n + tj.i0) * width;
    s.row01 = (tk.i0 * g.sz.mu_n + tj.i1) * width;
    s.row10 = (tk.i1 * g.sz.mu_n + tj.i0) * width;
    s.row11 = (tk.i1 * g.sz.mu_n + tj.i1) * width;
    s.w11 = tk.w * tj.w;
    s.w10 = tk.w - s.w11;
    s.w01 = tj.w - s.w11;
    s.w00 = 1.0f - tk.w - tj.w + s.w11;
    s.q = s.d_min = s.x_scale = s.cos_h = s.inv_sun_w = s.dens_r = s.dens_m = s.wy = 0.f;
    s.y0 = s.y1 = 0;
  } else {
    s.row00 = s.row01 = s.row10 = s.row11 = 0;
    s.w00 = s.w01 = s.w10 = s.w11 = 0.f;
    const double d_min = g.top - r_i, d_max = rho_i + g.H;
    s.q = (float)((g.top - r_i) * (g.top + r_i));
    s.d_min = (float)d_min;
    s.x_scale = (float)((g.sz.t_w - 1) / (d_max - d_min));
    const double sin_h = g.bottom / r_i;
    s.cos_h = (float)(-sqrt(d_pos(1.0 - sin_h * sin_h)));
    s.inv_sun_w = (float)(1.0 / (2.0 * sin_h * g.sun_angular_radius));
    const double h = r_i - g.bottom;
    s.dens_r = (float)profile_density(g.profiles[0], h);
    s.dens_m = (float)profile_density(g.profiles[1], h);
    const Tap ty = make_tap(rho_i / g.H * (g.sz.t_h - 1), g.sz.t_h);
    s.y0 = ty.i0; s.y1 = ty.i1; s.wy = ty.w;
  }
  *out = s;
  // GetTransmittance(r, mu, d, hit) (functions.glsl:493-519)
  double xa, ya, xb, yb;
  if (hit) {
    transmittance_xy(g, r_i, -mu_i, &xa, &ya);
    transmittance_xy(g, r, -mu, &xb, &yb);
  } else {
    transmittance_xy(g, r, mu, &xa, &ya);
    transmittance_xy(g, r_i, mu_i, &xb, &yb);
  }
  const Tap ax = make_tap(xa, g.sz.t_w), ay = make_tap(ya, g.sz.t_h);
  const Tap bx = make_tap(xb, g.sz.t_w), by = make_tap(yb, g.sz.t_h);
  const double w = ((i == 0 || i == PAS_RAY_SAMPLES) ? 0.5 : 1.0) * dx;
#pragma unroll 1
  for (int c = 0; c < NC; ++c) {
    const double t = fmin(fetch_t(T, CP, c, g.sz.t_w, ax, ay) / fetch_t(T, CP, c, g.sz.t_w, bx, by), 1.0);
    Tw[c] = (float)(t * w);
  }
#pragma unroll
  for (int c = NC; c < CP; ++c) Tw[c] = 0.f;
}

// RGBA store / accumulate into a final table (fp32 or fp16 texels).
__device__ __forceinline__ void final_rgba(void* base, size_t texel, float4 v, int half, bool add) {
  if (half) {
    __half2* p = reinterpret_cast<__half2*>(base) + 2 * texel;
    if (add) {
      const float2 a = __half22float2(p[0]), b = __half22float2(p[1]);
      v.x += a.x; v.y += a.y; v.z += b.x; v.w += b.y;
    }
    p[0] = __floats2half2_rn(v.x, v.y);
    p[1] = __floats2half2_rn(v.z, v.w);
  } else {
    float4* p = reinterpret_cast<float4*>(base) + texel;
    if (add) {
      const float4 a = *p;
      v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
    }
    *p = v;
  }
}

// Shared block prologue: ray of the block.
struct BlockRay {
  double r, rho, mu, d_end;
  bool hit;
};
__device__ __forceinline__ BlockRay block_ray(const PasGeometry& g, int k, int j) {
  BlockRay b;
  layer_radius(g, (k + 0.5) / g.sz.r_n, g.sz.r_n, &b.r, &b.rho);
  double r_mu;
  scattering_row_mu(g, b.r, b.rho, j, &b.mu, &r_mu, &b.hit);
  // DistanceToNearestAtmosphereBoundary (functions.glsl:680-687)
  b.d_end = b.hit ? dist_bottom(g, b.r, b.mu) : dist_top(g, b.r, b.mu);
  return b;
}

// Position (in float4 units) of vector q of texel x in a staged row of Q = CP / 4 vectors per texel.
// The XOR term spreads the same vector of 8 consecutive texels over the 8 16-byte bank groups.
template <int Q>
__device__ __forceinline__ int swz(int x, int q) {
  if (Q == 1) return x;
  constexpr int kShift = Q == 4 ? 1 : (Q == 2 ? 2 : 0);
  return x * Q + (q ^ ((x >> kShift) & (Q - 1)));
}

__device__ __forceinline__ float4 lerp4(float w, float4 a, float4 b) {
  return make_float4(fmaf(w, b.x - a.x, a.x), fmaf(w, b.y - a.y, a.y), fmaf(w, b.z - a.z, a.z),
                     fmaf(w, b.w - a.w, a.w));
}

// ---- multiple scattering ----------------------------------------------------------------------
template <int NC, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB)
multiple_scattering_kernel(const __grid_constant__ PasGeometry g,
                           const __grid_constant__ PasSpectrum sp, const float* __restrict__ T,
                           const float* __restrict__ dJ, float* __restrict__ dS, FinalTables fin,
                           int k_begin) {
  constexpr int CP = PAS_CHANNEL_PITCH(NC), Q = CP / 4;
  extern __shared__ __align__(16) float smem_dyn[];
  __shared__ RaySample sSample[kSamples];
  __shared__ __align__(16) float sTw[kSamples][CP];

  const int tid = threadIdx.x;
  const int j = blockIdx.x, k = k_begin + blockIdx.y;
  const int mu_n = g.sz.mu_n, nu_n = g.sz.nu_n, mu_s_n = g.sz.mu_s_n;
  const int width = nu_n * mu_s_n;
  float4* sRow = reinterpret_cast<float4*>(smem_dyn);  // [2][width * Q]
  const float4* dJ4 = reinterpret_cast<const float4*>(dJ);

  const BlockRay ray = block_ray(g, k, j);
  if (tid < kSamples) {
    ray_sample_setup<NC>(g, T, ray.r, ray.rho, ray.mu, ray.hit, ray.d_end, tid, true,
                         &sSample[tid], sTw[tid]);
  }

  // per-thread axes: mu_s (column) and nu (slab)
  const int x = tid;  // one block covers the whole row; threads >= width only help phase A
  const bool active = x < width;
  const int i_nu = active ? x / mu_s_n : 0, i_mu_s = active ? x % mu_s_n : 0;
  const double mu_s_d = scattering_col_mu_s(g, i_mu_s);
  const double nu_d = scattering_slab_nu(g, i_nu, ray.mu, mu_s_d);
  const float nu = (float)nu_d;
  const float r_mu_s = (float)(ray.r * mu_s_d);
  const float bottom = (float)g.bottom;
  // nu axis: slab index and lerp weight (functions.glsl:967-969), fixed along the ray
  const Tap tnu = make_tap((nu_d + 1.0) * 0.5 * (nu_n - 1), nu_n);
  const int slab0 = tnu.i0 * mu_s_n, slab1 = tnu.i1 * mu_s_n;
  const float wnu = tnu.w;
  MuSMap map;
  map.H2 = (float)(g.H * g.H);
  map.d_min = (float)(g.top - g.bottom);
  map.inv_range = (float)(1.0 / (g.H - (g.top - g.bottom)));
  map.inv_A = (float)(1.0 / g.mus_A);
  map.scale = (float)(mu_s_n - 1);

  float4 acc[Q];
#pragma unroll
  for (int q = 0; q < Q; ++q) acc[q] = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncthreads();

  if (ray.d_end > 0.0) {
    for (int i = 0; i < kSamples; ++i) {
      const RaySample s = sSample[i];
      float4* buf = sRow + (size_t)(i & 1) * width * Q;
      // stage: bilinear in (r, mu) applied to the whole row of texels. The row is walked as a flat
      // array of 16-byte vectors so that a warp reads 512 contiguous bytes per load.
      {
        const float4* p00 = dJ4 + (size_t)s.row00 * Q;
        const float4* p01 = dJ4 + (size_t)s.row01 * Q;
        const float4* p10 = dJ4 + (size_t)s.row10 * Q;
        const float4* p11 = dJ4 + (size_t)s.row11 * Q;
        const int nvec = width * Q;
#pragma unroll
        for (int it = 0; it < Q; ++it) {
          const int f = tid + it * (int)blockDim.x;  // blockDim.x >= width: Q rounds cover the row
          if (f < nvec) {
            const float4 a = __ldg(p00 + f), b = __ldg(p01 + f), c = __ldg(p10 + f), d = __ldg(p11 + f);
            float4 v;
            v.x = fmaf(s.w00, a.x, fmaf(s.w01, b.x, fmaf(s.w10, c.x, s.w11 * d.x)));
            v.y = fmaf(s.w00, a.y, fmaf(s.w01, b.y, fmaf(s.w10, c.y, s.w11 * d.y)));
            v.z = fmaf(s.w00, a.z, fmaf(s.w01, b.z, fmaf(s.w10, c.z, s.w11 * d.z)));
            v.w = fmaf(s.w00, a.w, fmaf(s.w01, b.w, fmaf(s.w10, c.w, s.w11 * d.w)));
            buf[swz<Q>(f / Q, f % Q)] = v;
          }
        }
      }
      __syncthreads();
      if (active) {
        // mu_s at the sample: (r mu_s + d nu) / r_i (functions.glsl:1314)
        const float mu_s_i = f_clamp(fmaf(s.d, nu, r_mu_s) * s.inv_r, -1.0f, 1.0f);
        const float xs = f_clamp(f_mu_s_texel_x(map, bottom * mu_s_i), 0.0f, map.scale);
        const Tap tm = make_tap_f(xs, mu_s_n);
        const float wm = tm.w;
        const int xa0 = slab0 + tm.i0, xa1 = slab0 + tm.i1, xb0 = slab1 + tm.i0, xb1 = slab1 + tm.i1;
        const float4* tw4 = reinterpret_cast<const float4*>(sTw[i]);
#pragma unroll
        for (int q = 0; q < Q; ++q) {
          const float4 va = lerp4(wm, buf[swz<Q>(xa0, q)], buf[swz<Q>(xa1, q)]);
          const float4 vb = lerp4(wm, buf[swz<Q>(xb0, q)], buf[swz<Q>(xb1, q)]);
          const float4 v = lerp4(wnu, va, vb);
          const float4 t = tw4[q];
          acc[q].x = fmaf(v.x, t.x, acc[q].x);
          acc[q].y = fmaf(v.y, t.y, acc[q].y);
          acc[q].z = fmaf(v.z, t.z, acc[q].z);
          acc[q].w = fmaf(v.w, t.w, acc[q].w);
        }
      }
    }
  }
  if (!active) return;
  const size_t texel = ((size_t)k * mu_n + j) * width + x;
  float4* out = reinterpret_cast<float4*>(dS) + texel * Q;
  float rgb[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int q = 0; q < Q; ++q) {
    out[q] = acc[q];
    const float v[4] = {acc[q].x, acc[q].y, acc[q].z, acc[q].w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int c = 4 * q + e;
      if (c < NC) {
#pragma unroll
        for (int a = 0; a < 3; ++a) rgb[a] = fmaf(sp.lum[a][c], v[e], rgb[a]);
      }
    }
  }
  // scattering += L . dS / RayleighPhaseFunction(nu) (model.cc:204-207), alpha += 0
  const float inv_pr = (float)(1.0 / rayleigh_phase(nu_d));
  final_rgba(fin.scattering, texel, make_float4(rgb[0] * inv_pr, rgb[1] * inv_pr, rgb[2] * inv_pr, 0.f),
             fin.half_precision, true);
}

// ---- single scattering ------------------------------------------------------------------------
template <int NC, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB)
single_scattering_kernel(const __grid_constant__ PasGeometry g,
                         const __grid_constant__ PasSpectrum sp, const float* __restrict__ T,
                         float* __restrict__ dR, float* __restrict__ dM, FinalTables fin,
                         int k_begin) {
  constexpr int CP = PAS_CHANNEL_PITCH(NC), Q = CP / 4;
  extern __shared__ __align__(16) float smem_dyn[];
  __shared__ RaySample sSample[kSamples];
  __shared__ __align__(16) float sTw[kSamples][CP];

  const int tid = threadIdx.x;
  const int j = blockIdx.x, k = k_begin + blockIdx.y;
  const int mu_n = g.sz.mu_n, nu_n = g.sz.nu_n, mu_s_n = g.sz.mu_s_n, t_w = g.sz.t_w;
  const int width = nu_n * mu_s_n;
  float4* sRow = reinterpret_cast<float4*>(smem_dyn);  // [2][t_w * Q]
  const float4* T4 = reinterpret_cast<const float4*>(T);

  const BlockRay ray = block_ray(g, k, j);
  if (tid < kSamples) {
    ray_sample_setup<NC>(g, T, ray.r, ray.rho, ray.mu, ray.hit, ray.d_end, tid, false,
                         &sSample[tid], sTw[tid]);
  }
  const int x = tid;
  const bool active = x < width;
  const int i_nu = active ? x / mu_s_n : 0, i_mu_s = active ? x % mu_s_n : 0;
  const double mu_s_d = scattering_col_mu_s(g, i_mu_s);
  const double nu_d = scattering_slab_nu(g, i_nu, ray.mu, mu_s_d);
  const float nu = (float)nu_d;
  const float r_mu_s = (float)(ray.r * mu_s_d);
  const float x_max = (float)(t_w - 1);

  float4 accR[Q], accM[Q];
#pragma unroll
  for (int q = 0; q < Q; ++q) accR[q] = accM[q] = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncthreads();

  if (ray.d_end > 0.0) {
    for (int i = 0; i < kSamples; ++i) {
      const RaySample s = sSample[i];
      float4* buf = sRow + (size_t)(i & 1) * t_w * Q;
      // stage the transmittance row at r_i (lerp of the two bracketing rows), flat 16-byte vectors
      {
        const float4* pa = T4 + (size_t)s.y0 * t_w * Q;
        const float4* pb = T4 + (size_t)s.y1 * t_w * Q;
        for (int f = tid; f < t_w * Q; f += blockDim.x) {
          buf[swz<Q>(f / Q, f % Q)] = lerp4(s.wy, __ldg(pa + f), __ldg(pb + f));
        }
      }
      __syncthreads();
      if (active) {
        // sun direction at the sample: r_i mu_s_i = r mu_s + d nu (functions.glsl:657)
        const float r_i = f_rcp(s.inv_r);
        const float p = f_clamp(fmaf(s.d, nu, r_mu_s), -r_i, r_i);
        // GetTransmittanceToSun (functions.glsl:552-563): table x from the distance to the top
        const float xt = f_clamp((f_dist_top(p, s.q) - s.d_min) * s.x_scale, 0.0f, x_max);
        const Tap tu = make_tap_f(xt, t_w);
        const float mu_s_i = p * s.inv_r;
        const float sm = f_sat(fmaf(mu_s_i - s.cos_h, s.inv_sun_w, 0.5f));
        const float vis = sm * sm * fmaf(-2.0f, sm, 3.0f);
        const float wr = vis * s.dens_r, wm = vis * s.dens_m;
        const float4* tw4 = reinterpret_cast<const float4*>(sTw[i]);
#pragma unroll
        for (int q = 0; q < Q; ++q) {
          const float4 tv = lerp4(tu.w, buf[swz<Q>(tu.i0, q)], buf[swz<Q>(tu.i1, q)]);
          const float4 t = tw4[q];
          const float vx = tv.x * t.x, vy = tv.y * t.y, vz = tv.z * t.z, vw = tv.w * t.w;
          accR[q].x = fmaf(vx, wr, accR[q].x); accM[q].x = fmaf(vx, wm, accM[q].x);
          accR[q].y = fmaf(vy, wr, accR[q].y); accM[q].y = fmaf(vy, wm, accM[q].y);
          accR[q].z = fmaf(vz, wr, accR[q].z); accM[q].z = fmaf(vz, wm, accM[q].z);
          accR[q].w = fmaf(vw, wr, accR[q].w); accM[q].w = fmaf(vw, wm, accM[q].w);
        }
      }
    }
  }
  if (!active) return;
  const size_t texel = ((size_t)k * mu_n + j) * width + x;
  float4* outR = reinterpret_cast<float4*>(dR) + texel * Q;
  float4* outM = reinterpret_cast<float4*>(dM) + texel * Q;
  float rgb[3] = {0.f, 0.f, 0.f}, mie[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int q = 0; q < Q; ++q) {
    float vr[4] = {accR[q].x, accR[q].y, accR[q].z, accR[q].w};
    float vm[4] = {accM[q].x, accM[q].y, accM[q].z, accM[q].w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int c = 4 * q + e;
      if (c < NC) {
        // functions.glsl:727-729 (dx is folded into sTw)
        vr[e] *= (float)(sp.solar[c] * sp.beta_r[c]);
        vm[e] *= (float)(sp.solar[c] * sp.beta_m_sca[c]);
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          rgb[a] = fmaf(sp.lum[a][c], vr[e], rgb[a]);
          mie[a] = fmaf(sp.lum[a][c], vm[e], mie[a]);
        }
      } else {
        vr[e] = vm[e] = 0.f;
      }
    }
    outR[q] = make_float4(vr[0], vr[1], vr[2], vr[3]);
    outM[q] = make_float4(vm[0], vm[1], vm[2], vm[3]);
  }
  // scattering = (L.dR, (L.dM).r), single_mie = L.dM (model.cc:151-156); blended when accumulating
  final_rgba(fin.scattering, texel, make_float4(rgb[0], rgb[1], rgb[2], mie[0]), fin.half_precision,
             fin.accumulate != 0);
  if (fin.single_mie != nullptr) {
    final_rgba(fin.single_mie, texel, make_float4(mie[0], mie[1], mie[2], 1.0f), fin.half_precision,
               fin.accumulate != 0);
  }
}

inline int round_up32(int v) { return (v + 31) / 32 * 32; }

// Rows of up to 256 texels (the reference's 8 x 32) run with 256-thread blocks and a register
// budget that keeps the 128-bit corner loads of a whole texel in flight; wider rows (up to 1024)
// fall back to one big block per row.
int tuning(const char* name, int dflt) {
  const char* e = getenv(name);
  return e != nullptr ? atoi(e) : dflt;
}

template <typename Kern>
cudaError_t prepare(Kern kern, size_t dyn) {
  if (dyn > 32 * 1024) {
    return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
  }
  return cudaSuccess;
}

template <int NC>
cudaError_t launch_multiple_nc(const PasGeometry& g, const PasSpectrum& s, const float* T,
                               const float* dJ, float* dS, FinalTables fin, int k_begin, int k_end,
                               cudaStream_t stream) {
  constexpr int CP = PAS_CHANNEL_PITCH(NC);
  const int width = g.sz.nu_n * g.sz.mu_s_n;
  if (width > 1024) return cudaErrorInvalidValue;
  const int threads = round_up32(width < kSamples ? kSamples : width);
  const size_t dyn = (size_t)2 * CP * width * sizeof(float);
  const dim3 grid(g.sz.mu_n, k_end - k_begin);
  cudaError_t e;
  if (threads <= 256) {
    static const int blocks = tuning("PAS_MS_BLOCKS", 3);
#define PAS_LAUNCH(B)                                                                         \
    {                                                                                         \
      auto kern = multiple_scattering_kernel<NC, 256, B>;                                     \
      if ((e = prepare(kern, dyn)) != cudaSuccess) return e;                                  \
      kern<<<grid, threads, dyn, stream>>>(g, s, T, dJ, dS, fin, k_begin);                    \
    }
    if (blocks == 2) PAS_LAUNCH(2) else if (blocks == 4) PAS_LAUNCH(4) else PAS_LAUNCH(3)
#undef PAS_LAUNCH
  } else {
    auto kern = multiple_scattering_kernel<NC, 1024, 1>;
    if ((e = prepare(kern, dyn)) != cudaSuccess) return e;
    kern<<<grid, threads, dyn, stream>>>(g, s, T, dJ, dS, fin, k_begin);
  }
  return cudaGetLastError();
}

template <int NC>
cudaError_t launch_single_nc(const PasGeometry& g, const PasSpectrum& s, const float* T, float* dR,
                             float* dM, FinalTables fin, int k_begin, int k_end,
                             cudaStream_t stream) {
  constexpr int CP = PAS_CHANNEL_PITCH(NC);
  const int width = g.sz.nu_n * g.sz.mu_s_n;
  if (width > 1024) return cudaErrorInvalidValue;
  const int threads = round_up32(width < kSamples ? kSamples : width);
  const size_t dyn = (size_t)2 * CP * g.sz.t_w * sizeof(float);
  const dim3 grid(g.sz.mu_n, k_end - k_begin);
  cudaError_t e;
  if (threads <= 256) {
    static const int blocks = tuning("PAS_SS_BLOCKS", 3);
#define PAS_LAUNCH(B)                                                                         \
    {                                                                                         \
      auto kern = single_scattering_kernel<NC, 256, B>;                                       \
      if ((e = prepare(kern, dyn)) != cudaSuccess) return e;                                  \
      kern<<<grid, threads, dyn, stream>>>(g, s, T, dR, dM, fin, k_begin);                    \
    }
    if (blocks == 2) PAS_LAUNCH(2) else if (blocks == 4) PAS_LAUNCH(4) else PAS_LAUNCH(3)
#undef PAS_LAUNCH
  } else {
    auto kern = single_scattering_kernel<NC, 1024, 1>;
    if ((e = prepare(kern, dyn)) != cudaSuccess) return e;
    kern<<<grid, threads, dyn, stream>>>(g, s, T, dR, dM, fin, k_begin);
  }
  return cudaGetLastError();
}

}  // namespace

cudaError_t launch_multiple_scattering(const PasGeometry& g, const PasSpectrum& s, const float* T,
                                       const float* dJ, float* dS, FinalTables fin, int k_begin,
                                       int k_end, cudaStream_t stream) {
  switch (s.nc) {
#define PAS_CASE(N) \
  case N: return launch_multiple_nc<N>(g, s, T, dJ, dS, fin, k_begin, k_end, stream);
    PAS_CASE(1) PAS_CASE(2) PAS_CASE(3) PAS_CASE(4) PAS_CASE(8) PAS_CASE(15) PAS_CASE(16)
#undef PAS_CASE
    default: return cudaErrorInvalidValue;
  }
}

cudaError_t launch_single_scattering(const PasGeometry& g, const PasSpectrum& s, const float* T,
                                     float* dR, float* dM, FinalTables fin, int k_begin, int k_end,
                                     cudaStream_t stream) {
  switch (s.nc) {
#define PAS_CASE(N) \
  case N: return launch_single_nc<N>(g, s, T, dR, dM, fin, k_begin, k_end, stream);
    PAS_CASE(1) PAS_CASE(2) PAS_CASE(3) PAS_CASE(4) PAS_CASE(8) PAS_CASE(15) PAS_CASE(16)
#undef PAS_CASE
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace pas

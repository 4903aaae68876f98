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
//   stage   (all threads): the shared-axis interpolation is applied ONCE to a whole table row,
//           coalesced from L2 into shared memory (multiple scattering: 4 (layer,row) corners ->
//           1 row of nu*mu_s values per channel; single scattering: 2 transmittance rows -> 1);
//   consume (per thread, fp32): only the thread-dependent axes remain (mu_s and nu, resp. the
//           sun-direction mu), i.e. 2-4 shared-memory reads per channel instead of 16 L2 reads.
#include "pas_kernels.h"
#include "pas_physics.cuh"

namespace pas {
namespace {

constexpr int kSamples = PAS_RAY_SAMPLES + 1;

__device__ __forceinline__ double fetch_t(const float* __restrict__ Tc, int w, const Tap& tx,
                                          const Tap& ty) {
  const double a = Tc[tx.i0 + w * ty.i0], b = Tc[tx.i1 + w * ty.i0];
  const double c = Tc[tx.i0 + w * ty.i1], d = Tc[tx.i1 + w * ty.i1];
  const double wx = tx.w, wy = ty.w;
  return a * ((1.0 - wx) * (1.0 - wy)) + b * (wx * (1.0 - wy)) + c * ((1.0 - wx) * wy) + d * (wx * wy);
}

// Shared per-sample record (written in phase A).
struct RaySample {
  float d;          // distance along the ray
  float inv_r;      // 1 / r_i
  // multiple scattering: footprint of (r_i, mu_i) in the source table
  int k0, k1, j0, j1;
  float wk, wj;
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
    s.k0 = tk.i0; s.k1 = tk.i1; s.wk = tk.w;
    s.j0 = tj.i0; s.j1 = tj.i1; s.wj = tj.w;
    s.q = s.d_min = s.x_scale = s.cos_h = s.inv_sun_w = s.dens_r = s.dens_m = s.wy = 0.f;
    s.y0 = s.y1 = 0;
  } else {
    s.k0 = s.k1 = s.j0 = s.j1 = 0;
    s.wk = s.wj = 0.f;
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
  const int nt = g.sz.t_w * g.sz.t_h;
#pragma unroll 1
  for (int c = 0; c < NC; ++c) {
    const float* Tc = T + (size_t)c * nt;
    const double t = fmin(fetch_t(Tc, g.sz.t_w, ax, ay) / fetch_t(Tc, g.sz.t_w, bx, by), 1.0);
    Tw[c] = (float)(t * w);
  }
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

// Shared block prologue: ray of the block and per-thread (mu_s, nu).
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

// ---- multiple scattering ----------------------------------------------------------------------
template <int NC>
__global__ void __launch_bounds__(1024)
multiple_scattering_kernel(const __grid_constant__ PasGeometry g,
                           const __grid_constant__ PasSpectrum sp, const float* __restrict__ T,
                           const float* __restrict__ dJ, float* __restrict__ dS, FinalTables fin,
                           int k_begin) {
  extern __shared__ __align__(16) float smem_dyn[];
  __shared__ RaySample sSample[kSamples];
  __shared__ float sTw[kSamples][NC];

  const int tid = threadIdx.x;
  const int j = blockIdx.x, k = k_begin + blockIdx.y;
  const int mu_n = g.sz.mu_n, nu_n = g.sz.nu_n, mu_s_n = g.sz.mu_s_n;
  const int width = nu_n * mu_s_n;
  const size_t row_stride = width;
  const size_t layer_stride = (size_t)width * mu_n;
  const size_t plane = layer_stride * g.sz.r_n;
  float* sRow = smem_dyn;  // [2][NC][width]

  const BlockRay ray = block_ray(g, k, j);
  if (tid < kSamples) {
    ray_sample_setup<NC>(g, T, ray.r, ray.rho, ray.mu, ray.hit, ray.d_end, tid, true,
                         &sSample[tid], sTw[tid]);
  }

  // per-thread axes: mu_s (column) and nu (slab)
  const int x = tid;  // one block covers the whole row; threads >= width only help staging
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

  float acc[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c) acc[c] = 0.f;
  __syncthreads();

  if (ray.d_end > 0.0) {
    for (int i = 0; i < kSamples; ++i) {
      const RaySample s = sSample[i];
      float* buf = sRow + (size_t)(i & 1) * NC * width;
      // stage: bilinear in (r, mu) applied to whole rows, coalesced
      {
        const float w11 = s.wk * s.wj, w10 = s.wk - w11, w01 = s.wj - w11;
        const float w00 = 1.0f - s.wk - s.wj + w11;
        const float* p00 = dJ + s.k0 * layer_stride + s.j0 * row_stride;
        const float* p01 = dJ + s.k0 * layer_stride + s.j1 * row_stride;
        const float* p10 = dJ + s.k1 * layer_stride + s.j0 * row_stride;
        const float* p11 = dJ + s.k1 * layer_stride + s.j1 * row_stride;
        for (int xx = tid; xx < width; xx += blockDim.x) {
#pragma unroll
          for (int c = 0; c < NC; ++c) {
            const size_t o = (size_t)c * plane + xx;
            buf[c * width + xx] =
                fmaf(w00, p00[o], fmaf(w01, p01[o], fmaf(w10, p10[o], w11 * p11[o])));
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
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          const float* b = buf + c * width;
          const float a0 = b[slab0 + tm.i0], a1 = b[slab0 + tm.i1];
          const float b0 = b[slab1 + tm.i0], b1 = b[slab1 + tm.i1];
          const float va = fmaf(wm, a1 - a0, a0), vb = fmaf(wm, b1 - b0, b0);
          acc[c] = fmaf(fmaf(wnu, vb - va, va), sTw[i][c], acc[c]);
        }
      }
    }
  }
  if (!active) return;
  const size_t texel = (size_t)k * layer_stride + (size_t)j * row_stride + x;
  float rgb[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    dS[(size_t)c * plane + texel] = acc[c];
#pragma unroll
    for (int a = 0; a < 3; ++a) rgb[a] = fmaf(sp.lum[a][c], acc[c], rgb[a]);
  }
  // scattering += L . dS / RayleighPhaseFunction(nu) (model.cc:204-207), alpha += 0
  const float inv_pr = (float)(1.0 / rayleigh_phase(nu_d));
  final_rgba(fin.scattering, texel, make_float4(rgb[0] * inv_pr, rgb[1] * inv_pr, rgb[2] * inv_pr, 0.f),
             fin.half_precision, true);
}

// ---- single scattering ------------------------------------------------------------------------
template <int NC>
__global__ void __launch_bounds__(1024)
single_scattering_kernel(const __grid_constant__ PasGeometry g,
                         const __grid_constant__ PasSpectrum sp, const float* __restrict__ T,
                         float* __restrict__ dR, float* __restrict__ dM, FinalTables fin,
                         int k_begin) {
  extern __shared__ __align__(16) float smem_dyn[];
  __shared__ RaySample sSample[kSamples];
  __shared__ float sTw[kSamples][NC];

  const int tid = threadIdx.x;
  const int j = blockIdx.x, k = k_begin + blockIdx.y;
  const int mu_n = g.sz.mu_n, nu_n = g.sz.nu_n, mu_s_n = g.sz.mu_s_n, t_w = g.sz.t_w;
  const int width = nu_n * mu_s_n;
  const size_t layer_stride = (size_t)width * mu_n;
  const size_t plane = layer_stride * g.sz.r_n;
  const size_t t_plane = (size_t)t_w * g.sz.t_h;
  float* sRow = smem_dyn;  // [2][NC][t_w]

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

  float accR[NC], accM[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c) accR[c] = accM[c] = 0.f;
  __syncthreads();

  if (ray.d_end > 0.0) {
    for (int i = 0; i < kSamples; ++i) {
      const RaySample s = sSample[i];
      float* buf = sRow + (size_t)(i & 1) * NC * t_w;
      // stage the transmittance row at r_i (lerp of the two bracketing rows)
      for (int u = tid; u < t_w; u += blockDim.x) {
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          const float* Tc = T + (size_t)c * t_plane;
          const float a = Tc[s.y0 * t_w + u], b = Tc[s.y1 * t_w + u];
          buf[c * t_w + u] = fmaf(s.wy, b - a, a);
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
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          const float* b = buf + c * t_w;
          const float t0 = b[tu.i0], t1 = b[tu.i1];
          const float tv = fmaf(tu.w, t1 - t0, t0) * sTw[i][c];
          accR[c] = fmaf(tv, wr, accR[c]);
          accM[c] = fmaf(tv, wm, accM[c]);
        }
      }
    }
  }
  if (!active) return;
  const size_t texel = (size_t)k * layer_stride + (size_t)j * width + x;
  float rgb[3] = {0.f, 0.f, 0.f}, mie[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    // functions.glsl:727-729 (dx is folded into sTw)
    const float ray_c = accR[c] * (float)(sp.solar[c] * sp.beta_r[c]);
    const float mie_c = accM[c] * (float)(sp.solar[c] * sp.beta_m_sca[c]);
    dR[(size_t)c * plane + texel] = ray_c;
    dM[(size_t)c * plane + texel] = mie_c;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      rgb[a] = fmaf(sp.lum[a][c], ray_c, rgb[a]);
      mie[a] = fmaf(sp.lum[a][c], mie_c, mie[a]);
    }
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

template <int NC>
cudaError_t launch_multiple_nc(const PasGeometry& g, const PasSpectrum& s, const float* T,
                               const float* dJ, float* dS, FinalTables fin, int k_begin, int k_end,
                               cudaStream_t stream) {
  const int width = g.sz.nu_n * g.sz.mu_s_n;
  if (width > 1024) return cudaErrorInvalidValue;
  const int threads = round_up32(width < kSamples ? kSamples : width);
  const size_t dyn = (size_t)2 * NC * width * sizeof(float);
  auto kern = multiple_scattering_kernel<NC>;
  if (dyn > 32 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
    if (e != cudaSuccess) return e;
  }
  kern<<<dim3(g.sz.mu_n, k_end - k_begin), threads, dyn, stream>>>(g, s, T, dJ, dS, fin, k_begin);
  return cudaGetLastError();
}

template <int NC>
cudaError_t launch_single_nc(const PasGeometry& g, const PasSpectrum& s, const float* T, float* dR,
                             float* dM, FinalTables fin, int k_begin, int k_end,
                             cudaStream_t stream) {
  const int width = g.sz.nu_n * g.sz.mu_s_n;
  if (width > 1024) return cudaErrorInvalidValue;
  const int threads = round_up32(width < kSamples ? kSamples : width);
  const size_t dyn = (size_t)2 * NC * g.sz.t_w * sizeof(float);
  auto kern = single_scattering_kernel<NC>;
  if (dyn > 32 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
    if (e != cudaSuccess) return e;
  }
  kern<<<dim3(g.sz.mu_n, k_end - k_begin), threads, dyn, stream>>>(g, s, T, dR, dM, fin, k_begin);
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

// Indirect irradiance pass (ComputeIndirectIrradianceTexture, atmosphere/functions.glsl:1477-1511,
// 1573-1586) with the fused accumulation E += L . dE (atmosphere/model.cc:176-190).
//
// One block per irradiance texel (r, mu_s). All 16 x 64 directions of the hemisphere integral look
// the radiance table up at the texel's own r and mu_s, and at mu = cos(theta_j): only nu varies
// with the azimuth. The block therefore first reduces the three shared axes once per polar ring
// (8 corners -> one value per (ring, channel, nu slab) in shared memory), then every thread
// evaluates its directions with a 1-D nu interpolation. The 64 azimuths pair up (phi, -phi) with
// equal nu, so 32 cosines x weight 2 are evaluated.
#include "pas_kernels.h"
#include "pas_physics.cuh"

namespace pas {
namespace {

constexpr int kThreads = 256;

template <int NC, bool ORDER1>
__global__ void __launch_bounds__(kThreads)
indirect_irradiance_kernel(const __grid_constant__ PasGeometry g,
                           const __grid_constant__ PasSpectrum sp, const float* __restrict__ tabA,
                           const float* __restrict__ tabB, float* __restrict__ dE, FinalTables fin,
                           int j_begin, int k_begin, int k_end, int k_stride) {
  constexpr int NT = ORDER1 ? 2 : 1;
  constexpr int CP = PAS_CHANNEL_PITCH(NC);
  __shared__ float sV[NT][PAS_IRR_THETA][NC][PAS_MAX_NU];
  __shared__ float sRed[kThreads / 32][NC];

  const int tid = threadIdx.x;
  const int i = blockIdx.x, j = j_begin + blockIdx.y;
  const int mu_n = g.sz.mu_n, nu_n = g.sz.nu_n, mu_s_n = g.sz.mu_s_n, r_n = g.sz.r_n;
  const int width = nu_n * mu_s_n;
  const size_t layer_stride = (size_t)width * mu_n;

  // texel -> (r, mu_s), functions.glsl:1539-1548
  const double r = g.bottom + unit_from_coord((j + 0.5) / g.sz.e_h, g.sz.e_h) * (g.top - g.bottom);
  const double mu_s = d_clamp(2.0 * unit_from_coord((i + 0.5) / g.sz.e_w, g.sz.e_w) - 1.0, -1.0, 1.0);
  const double rho = sqrt(d_pos(r * r - g.bottom * g.bottom));
  const Tap tk = make_tap(rho / g.H * (r_n - 1), r_n);
  const Tap ts = make_tap(scattering_x_from_mu_s(g, mu_s), mu_s_n);

  // per-ring and per-azimuth constants, once per block (fp64 trigonometry, then fp32)
  __shared__ Tap sTj[PAS_IRR_THETA];
  __shared__ float sSinT[PAS_IRR_THETA], sCosT[PAS_IRR_THETA], sCosP[PAS_IRR_PHI / 2];
  if (tid < PAS_IRR_THETA) {
    const double theta = (tid + 0.5) * (kPi / (2 * PAS_IRR_THETA));
    sTj[tid] = make_tap(scattering_y_from_mu(g, r, rho, cos(theta), false), mu_n);
    sSinT[tid] = (float)sin(theta);
    sCosT[tid] = (float)cos(theta);
  } else if (tid >= 32 && tid < 32 + PAS_IRR_PHI / 2) {
    const double phi = (tid - 32 + 0.5) * (kPi / (PAS_IRR_PHI / 2));
    sCosP[tid - 32] = (float)cos(phi);
  }
  __syncthreads();

  // stage: reduce (r, mu, mu_s) once per (ring, channel, slab); unrolled so that the 8 corner loads of
  // several items are in flight together (the kernel is one wave of latency-bound blocks)
#pragma unroll 4
  for (int idx = tid; idx < NT * PAS_IRR_THETA * NC * nu_n; idx += kThreads) {
    const int c = idx % NC, s = (idx / NC) % nu_n, l = (idx / (nu_n * NC)) % PAS_IRR_THETA;
    const int t = idx / (nu_n * NC * PAS_IRR_THETA);
    const Tap tj = sTj[l];
    const float* p = (t == 0 ? tabA : tabB) + c;  // interleaved: tab[texel * CP + c]
    float v = 0.f;
#pragma unroll
    for (int corner = 0; corner < 8; ++corner) {
      const int kk = (corner & 4) ? tk.i1 : tk.i0;
      const int jj = (corner & 2) ? tj.i1 : tj.i0;
      const int ii = (corner & 1) ? ts.i1 : ts.i0;
      const float w = ((corner & 4) ? tk.w : 1.f - tk.w) * ((corner & 2) ? tj.w : 1.f - tj.w) *
                      ((corner & 1) ? ts.w : 1.f - ts.w);
      // r-slab sharding: layers owned by other ranks contribute through their partial sums
      if (kk < k_begin || kk >= k_end || (kk - k_begin) % k_stride != 0) continue;
      v = fmaf(w, p[(kk * layer_stride + (size_t)jj * width + s * mu_s_n + ii) * CP], v);
    }
    sV[t][l][c][s] = v;
  }
  __syncthreads();

  // directions: 16 rings x 32 cosines (each standing for +-phi)
  const float sx = (float)sqrt(1.0 - mu_s * mu_s), sz = (float)mu_s;
  const float scale = 0.5f * (float)(nu_n - 1);
  const float mie_g = (float)g.mie_g;
  const float g2p1 = 1.0f + mie_g * mie_g, m2g = -2.0f * mie_g;
  const float kR1 = (float)(3.0 / (16.0 * kPi)), kM1 = (float)mie_phase_k(g.mie_g);
  float acc[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c) acc[c] = 0.f;
  for (int d = tid; d < PAS_IRR_THETA * (PAS_IRR_PHI / 2); d += kThreads) {
    const int l = d / (PAS_IRR_PHI / 2), m = d % (PAS_IRR_PHI / 2);
    const float st = sSinT[l], ct = sCosT[l], cp = sCosP[m];
    // domega * omega.z, both azimuth signs (functions.glsl:1500-1507)
    const float w = 2.0f * ct * st * (float)((kPi / (2 * PAS_IRR_THETA)) * (kPi / (PAS_IRR_PHI / 2)));
    const float nu = fmaf(cp * st, sx, ct * sz);
    const float xn = f_clamp(fmaf(nu, scale, scale), 0.f, 2.0f * scale);
    const Tap tn = make_tap_f(xn, nu_n);
    float wa = w, wb = 0.f;
    if (ORDER1) {
      const float q = fmaf(nu, nu, 1.0f);
      const float rs = f_rsqrt(fmaf(m2g, nu, g2p1));
      wa = w * kR1 * q;
      wb = w * (kM1 * q) * (rs * rs) * rs;
    }
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const float a0 = sV[0][l][c][tn.i0], a1 = sV[0][l][c][tn.i1];
      acc[c] = fmaf(fmaf(tn.w, a1 - a0, a0), wa, acc[c]);
      if (ORDER1) {
        const float b0 = sV[NT - 1][l][c][tn.i0], b1 = sV[NT - 1][l][c][tn.i1];
        acc[c] = fmaf(fmaf(tn.w, b1 - b0, b0), wb, acc[c]);
      }
    }
  }
  // block reduction
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    float v = acc[c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((tid & 31) == 0) sRed[tid >> 5][c] = v;
  }
  __syncthreads();
  if (tid == 0) {
    const int n = g.sz.e_w * g.sz.e_h;
    const int t = j * g.sz.e_w + i;
    float rgb[3] = {0.f, 0.f, 0.f};
    for (int c = 0; c < NC; ++c) {
      float v = 0.f;
      for (int w = 0; w < kThreads / 32; ++w) v += sRed[w][c];
      dE[(size_t)c * n + t] = v;
      for (int a = 0; a < 3; ++a) rgb[a] = fmaf(sp.lum[a][c], v, rgb[a]);
    }
    if (fin.irradiance != nullptr) {
      float4* p = reinterpret_cast<float4*>(fin.irradiance) + t;
      float4 e = *p;
      e.x += rgb[0]; e.y += rgb[1]; e.z += rgb[2];
      *p = e;
    }
  }
}

template <int NC>
cudaError_t launch_nc(const PasGeometry& g, const PasSpectrum& s, const float* dR, const float* dM,
                      const float* dS, int order, float* dE, FinalTables fin, int j_begin,
                      int j_end, LayerSet layers, cudaStream_t stream) {
  dim3 grid(g.sz.e_w, j_end - j_begin);
  if (order == 1) {
    indirect_irradiance_kernel<NC, true><<<grid, kThreads, 0, stream>>>(g, s, dR, dM, dE, fin, j_begin,
                                                                        layers.begin, layers.end, layers.stride);
  } else {
    indirect_irradiance_kernel<NC, false><<<grid, kThreads, 0, stream>>>(g, s, dS, nullptr, dE, fin,
                                                                         j_begin, layers.begin, layers.end, layers.stride);
  }
  return cudaGetLastError();
}

}  // namespace

cudaError_t launch_indirect_irradiance(const PasGeometry& g, const PasSpectrum& s, const float* dR,
                                       const float* dM, const float* dS, int order, float* dE,
                                       FinalTables fin, int j_begin, int j_end, LayerSet layers,
                                       cudaStream_t stream) {
  if (g.sz.nu_n > PAS_MAX_NU) return cudaErrorInvalidValue;
  switch (s.nc) {
#define PAS_CASE(N) \
  case N: return launch_nc<N>(g, s, dR, dM, dS, order, dE, fin, j_begin, j_end, layers, stream);
    PAS_CASE(3) PAS_CASE(4) PAS_CASE(8) PAS_CASE(15) PAS_CASE(16)
#undef PAS_CASE
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace pas

// Scattering density pass (ComputeScatteringDensityTexture, atmosphere/functions.glsl:1163-1260,
// 1348-1367): ~80-90 % of the reference's work.
//
// Formulation (DESIGN.md "density kernel"): every table lookup of this pass happens at the output
// texel's own r and mu_s, so the 4-D fetch of functions.glsl:958-976 degenerates to a bilinear
// fetch in (mu, nu) at a fixed (layer k, column i_mu_s); and the mu coordinate depends on
// (k, theta_l) only. A block therefore owns one (k, i_mu_s) pair and stages, per polar direction
// l, the mu-interpolated row A[l][c][0..NU) of the previous order's table(s) in shared memory.
// The remaining nu interpolation is a piecewise-linear function of nu1 = omega_s . omega_i with NU
// uniform knots,   L(x) = V[0] + sum_s (V[s+1] - V[s]) * sat(x - s),   x = (nu1 + 1)(NU - 1)/2,
// which is LINEAR in the table values. The thread (one output texel) therefore accumulates the
// channel-independent weights  W[s] = sum_m sat(x_m - s) * phase(nu2_m) * domega  over the 32
// azimuths in registers, and contracts them with the C channels once per l. The ground term
// (functions.glsl:1234-1240) is piecewise linear in nu1 too, because the cosine at the ground
// point is affine in nu1: (r mu_s + d_g nu1) / bottom; it is handled the same way over a small
// window of irradiance knots. Per direction the work is ~25-50 fp32 instructions independent of
// the channel count, instead of ~(36 + 16 C) flops.
//
// The azimuth samples come in pairs (phi, -phi) sharing cos(phi), hence nu2 and both phase
// functions (functions.glsl:1246-1256): the loop runs over 16 cosines x 2 signs of sin(phi).
#include "pas_kernels.h"
#include "pas_physics.cuh"

namespace pas {
namespace {

constexpr int kNG = 4;  // irradiance ramps handled per sweep of the azimuth loop

// cos/sin of phi_m = (m + 0.5) pi / 16, m = 0..15 (functions.glsl:1216); m' = 31 - m mirrors sin.
__device__ constexpr float kCosPhi[16] = {
    0.99518472f, 0.956940353f, 0.881921291f, 0.773010433f, 0.634393275f, 0.471396744f,
    0.290284663f, 0.0980171412f, -0.0980171412f, -0.290284663f, -0.471396744f, -0.634393275f,
    -0.773010433f, -0.881921291f, -0.956940353f, -0.99518472f};
__device__ constexpr float kSinPhi[16] = {
    0.0980171412f, 0.290284663f, 0.471396744f, 0.634393275f, 0.773010433f, 0.881921291f,
    0.956940353f, 0.99518472f, 0.99518472f, 0.956940353f, 0.881921291f, 0.773010433f,
    0.634393275f, 0.471396744f, 0.290284663f, 0.0980171412f};

struct DirConst {       // per (block, l), in shared memory
  float cos_t, sin_t;
  float g_a, g_b;       // ground cosine -> irradiance texel x: x = g_a + g_b * nu1
  int hit;
  int win_i0;           // first irradiance knot of the window
  int win_n;            // number of ramps needed (>= 0)
  int pad;
};

// Weights accumulated by one thread for one polar direction.
template <bool ORDER2, int NUM>
struct Weights {
  // Order >= 3: w[0] pairs with the Rayleigh coefficient, w[1] with the Mie one.
  // Order 2: w[0..3] = (table R, coef R), (table R, coef M), (table M, coef R), (table M, coef M).
  static constexpr int NW = ORDER2 ? 4 : 2;
  float base[NW];
  float ramp[NW][NUM - 1];
  float gbase[2];          // sum of plain phase weights (ground term base), coef R / coef M
  float gramp[2][kNG];
};

template <bool ORDER2, int NUM, bool HIT, bool TABLES>
__device__ __forceinline__ void azimuth_sweep(Weights<ORDER2, NUM>& W, float x0, float xc, float xs,
                                              float n0, float nc_, float v0, float vc, float vs,
                                              float gx0, float gxc, float gxs, float kR, float kM,
                                              float kR1, float kM1, float g2p1, float m2g) {
#pragma unroll
  for (int m = 0; m < 16; ++m) {
    const float c = kCosPhi[m], s = kSinPhi[m];
    // phase functions towards the view direction (functions.glsl:1246-1256), shared by +-phi
    const float nu2 = fmaf(nc_, c, n0);
    const float q = fmaf(nu2, nu2, 1.0f);
    const float pR = kR * q;
    const float rs = f_rsqrt(fmaf(m2g, nu2, g2p1));
    const float pM = (kM * q) * (rs * rs) * rs;
    if (TABLES || HIT) {
      W.gbase[0] += pR;
      W.gbase[1] += pM;
    }
    if (TABLES) {
      const float xb = fmaf(xc, c, x0);
      const float xp = fmaf(xs, s, xb), xm = fmaf(-xs, s, xb);
      if (!ORDER2) {
#pragma unroll
        for (int k = 0; k < NUM - 1; ++k) {
          const float cs = f_sat(xp - (float)k) + f_sat(xm - (float)k);
          W.ramp[0][k] = fmaf(cs, pR, W.ramp[0][k]);
          W.ramp[1][k] = fmaf(cs, pM, W.ramp[1][k]);
        }
      } else {
        // order 2: incident radiance = R * P_R(nu1) + M * P_M(nu1) (functions.glsl:995-1003)
        const float vb = fmaf(vc, c, v0);
        const float nu1p = fmaf(vs, s, vb), nu1m = fmaf(-vs, s, vb);
        const float q1p = fmaf(nu1p, nu1p, 1.0f), q1m = fmaf(nu1m, nu1m, 1.0f);
        const float rp = f_rsqrt(fmaf(m2g, nu1p, g2p1)), rm = f_rsqrt(fmaf(m2g, nu1m, g2p1));
        const float PRp = kR1 * q1p, PRm = kR1 * q1m;
        const float PMp = (kM1 * q1p) * (rp * rp) * rp, PMm = (kM1 * q1m) * (rm * rm) * rm;
        W.base[0] = fmaf(PRp + PRm, pR, W.base[0]);
        W.base[1] = fmaf(PRp + PRm, pM, W.base[1]);
        W.base[2] = fmaf(PMp + PMm, pR, W.base[2]);
        W.base[3] = fmaf(PMp + PMm, pM, W.base[3]);
#pragma unroll
        for (int k = 0; k < NUM - 1; ++k) {
          const float cp = f_sat(xp - (float)k), cm = f_sat(xm - (float)k);
          const float a = fmaf(cp, PRp, cm * PRm);  // table R weight
          const float b = fmaf(cp, PMp, cm * PMm);  // table M weight
          W.ramp[0][k] = fmaf(a, pR, W.ramp[0][k]);
          W.ramp[1][k] = fmaf(a, pM, W.ramp[1][k]);
          W.ramp[2][k] = fmaf(b, pR, W.ramp[2][k]);
          W.ramp[3][k] = fmaf(b, pM, W.ramp[3][k]);
        }
      }
    }
    if (HIT) {
      const float gb = fmaf(gxc, c, gx0);
      const float gp = fmaf(gxs, s, gb), gm = fmaf(-gxs, s, gb);
#pragma unroll
      for (int k = 0; k < kNG; ++k) {
        const float cs = f_sat(gp - (float)k) + f_sat(gm - (float)k);
        W.gramp[0][k] = fmaf(cs, pR, W.gramp[0][k]);
        W.gramp[1][k] = fmaf(cs, pM, W.gramp[1][k]);
      }
    }
  }
}

template <int NC, bool ORDER2, int NUM>
__global__ void __launch_bounds__(256, 2)
density_kernel(const __grid_constant__ PasGeometry g, const PasDensityDir* __restrict__ dirs,
               const float* __restrict__ G, const float* __restrict__ cRk,
               const float* __restrict__ cMk, const float* __restrict__ tabA,
               const float* __restrict__ tabB, const float* __restrict__ dE,
               float* __restrict__ dJ, int k_begin) {
  constexpr int NT = ORDER2 ? 2 : 1;  // tables staged
  constexpr int CP = PAS_CHANNEL_PITCH(NC);
  extern __shared__ __align__(16) float smem_dyn[];
  __shared__ __align__(16) float sA[NT][PAS_DIR_THETA][NC][NUM];
  __shared__ float sG[PAS_DIR_THETA][NC];
  __shared__ float sCR[NC], sCM[NC];
  __shared__ DirConst sDir[PAS_DIR_THETA];

  const int tid = threadIdx.x;
  const int i_mu_s = blockIdx.y;
  const int k = k_begin + blockIdx.z;
  const int mu_n = g.sz.mu_n, nu_n = g.sz.nu_n, mu_s_n = g.sz.mu_s_n, e_w = g.sz.e_w;
  const int width = nu_n * mu_s_n;
  const size_t layer = (size_t)k * mu_n * width;
  // dynamic shared memory: irradiance row 0 and its forward differences, zero padded
  const int e_pad = e_w + kNG + 1;
  float* sE0 = smem_dyn;              // [NC][e_pad]
  float* sDE = smem_dyn + NC * e_pad; // [NC][e_pad]

  double r, rho;
  layer_radius(g, (k + 0.5) / g.sz.r_n, g.sz.r_n, &r, &rho);
  const double mu_s_d = scattering_col_mu_s(g, i_mu_s);

  // ---- stage the block's tables --------------------------------------------------------------
  if (tid < PAS_DIR_THETA) {
    const PasDensityDir d = dirs[k * PAS_DIR_THETA + tid];
    DirConst dc;
    dc.cos_t = d.cos_t;
    dc.sin_t = d.sin_t;
    dc.hit = d.hit;
    // cosine at the ground point: (r mu_s + d_g nu1) / bottom  (|zenith r + omega_i d_g| = bottom,
    // functions.glsl:1234-1238); irradiance texel x = (cos/2 + 1/2)(e_w - 1) (functions.glsl:1530-1531)
    const double half = 0.5 * (e_w - 1);
    const double ga = (r * mu_s_d / g.bottom) * half + half;
    const double gb = (double)d.dg_over_b * half;
    dc.g_a = (float)ga;
    dc.g_b = (float)gb;
    // knots [i0, i1] cover every x the azimuth loop can produce (nu1 in [-1, 1]); x outside
    // [0, e_w - 1] is clamped by the saturating ramps exactly like the fetch's index clamp
    int i0 = (int)floor(ga - gb - 1e-3);
    i0 = i0 < 0 ? 0 : (i0 > e_w - 2 ? e_w - 2 : i0);
    int i1 = (int)ceil(ga + gb + 1e-3);
    i1 = i1 > e_w - 1 ? e_w - 1 : (i1 < i0 + 1 ? i0 + 1 : i1);
    dc.win_i0 = i0;
    dc.win_n = d.hit ? i1 - i0 : 0;
    dc.pad = 0;
    sDir[tid] = dc;
  }
  if (tid < NC) {
    sCR[tid] = cRk[k * PAS_MAX_CH + tid];
    sCM[tid] = cMk[k * PAS_MAX_CH + tid];
  }
  for (int idx = tid; idx < PAS_DIR_THETA * NC; idx += blockDim.x) {
    sG[idx / NC][idx % NC] = G[(size_t)(k * PAS_DIR_THETA + idx / NC) * PAS_MAX_CH + idx % NC];
  }
  for (int idx = tid; idx < NC * e_pad; idx += blockDim.x) {
    const int c = idx / e_pad, i = idx % e_pad;
    const float* row = dE + (size_t)c * e_w * g.sz.e_h;  // row 0: r = bottom
    sE0[idx] = i < e_w ? row[i] : 0.f;
    sDE[idx] = i < e_w - 1 ? row[i + 1] - row[i] : 0.f;
  }
  // mu-interpolated rows: value at slab s, then in place -> (V[0], D[0..NUM-2]). Consecutive
  // threads read the consecutive channels of one interleaved texel (64 B at 15 channels).
  for (int idx = tid; idx < NT * PAS_DIR_THETA * NUM * NC; idx += blockDim.x) {
    const int c = idx % NC, s = (idx / NC) % NUM, l = (idx / (NC * NUM)) % PAS_DIR_THETA;
    const int t = idx / (NC * NUM * PAS_DIR_THETA);
    float v = 0.f;
    if (s < nu_n) {
      const PasDensityDir d = dirs[k * PAS_DIR_THETA + l];
      const float* tab = (t == 0 ? tabA : tabB) + (layer + s * mu_s_n + i_mu_s) * CP + c;
      const float a = tab[(size_t)d.j0 * width * CP], b = tab[(size_t)d.j1 * width * CP];
      v = fmaf(d.w_row, b - a, a);
    }
    sA[t][l][c][s] = v;
  }
  __syncthreads();
  for (int row = tid; row < NT * PAS_DIR_THETA * NC; row += blockDim.x) {
    float* p = (&sA[0][0][0][0]) + row * NUM;
    float v[NUM];
#pragma unroll
    for (int s = 0; s < NUM; ++s) v[s] = p[s];
#pragma unroll
    for (int s = 0; s < NUM - 1; ++s) p[1 + s] = (s + 1 < nu_n) ? v[s + 1] - v[s] : 0.f;
  }
  __syncthreads();

  // ---- per-texel geometry (fp64, once) --------------------------------------------------------
  const int texel = blockIdx.x * blockDim.x + tid;
  if (texel >= mu_n * nu_n) return;
  const int j = texel / nu_n, i_nu = texel % nu_n;
  double mu_d, r_mu_d;
  bool hit_unused;
  scattering_row_mu(g, r, rho, j, &mu_d, &r_mu_d, &hit_unused);
  const double nu_d = scattering_slab_nu(g, i_nu, mu_d, mu_s_d);
  // omega = (sqrt(1 - mu^2), 0, mu), omega_s = (sx, sy, mu_s) (functions.glsl:1181-1185)
  const double wx_d = sqrt(1.0 - mu_d * mu_d);
  const double sx_d = wx_d == 0.0 ? 0.0 : (nu_d - mu_d * mu_s_d) / wx_d;
  const double sy_d = sqrt(d_pos(1.0 - sx_d * sx_d - mu_s_d * mu_s_d));
  const float wx = (float)wx_d, mu = (float)mu_d, sx = (float)sx_d, sy = (float)sy_d;
  const float mu_s = (float)mu_s_d;

  const float scale = 0.5f * (float)(nu_n - 1);
  const float mie_g = (float)g.mie_g;
  const float g2p1 = 1.0f + mie_g * mie_g, m2g = -2.0f * mie_g;
  const float kR1 = (float)(3.0 / (16.0 * kPi));
  const float kM1 = (float)mie_phase_k(g.mie_g);
  const float dtheta_dphi = (float)((kPi / PAS_DIR_THETA) * (kPi / PAS_DIR_THETA));

  float acc[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c) acc[c] = 0.f;

  for (int l = 0; l < PAS_DIR_THETA; ++l) {
    const DirConst dc = sDir[l];
    const float ct = dc.cos_t, st = dc.sin_t;
    const float domega = dtheta_dphi * st;  // functions.glsl:1219
    // nu1 = omega_s . omega_i = v0 + vc cos(phi) + vs sin(phi); x = (nu1 + 1) * scale
    const float v0 = mu_s * ct, vc = sx * st, vs = sy * st;
    const float x0 = fmaf(v0, scale, scale), xc = vc * scale, xs = vs * scale;
    // nu2 = omega . omega_i = n0 + nc cos(phi)
    const float n0 = mu * ct, ncf = wx * st;
    const float kR = kR1 * domega, kM = kM1 * domega;

    Weights<ORDER2, NUM> W;
#pragma unroll
    for (int a = 0; a < Weights<ORDER2, NUM>::NW; ++a) {
      W.base[a] = 0.f;
#pragma unroll
      for (int s = 0; s < NUM - 1; ++s) W.ramp[a][s] = 0.f;
    }
    W.gbase[0] = W.gbase[1] = 0.f;
#pragma unroll
    for (int t = 0; t < kNG; ++t) W.gramp[0][t] = W.gramp[1][t] = 0.f;

    // ground window in texel units relative to its first knot
    const float gx0 = fmaf(dc.g_b, v0, dc.g_a) - (float)dc.win_i0;
    const float gxc = dc.g_b * vc, gxs = dc.g_b * vs;
    if (dc.win_n > 0) {
      azimuth_sweep<ORDER2, NUM, true, true>(W, x0, xc, xs, n0, ncf, v0, vc, vs, gx0, gxc, gxs, kR,
                                             kM, kR1, kM1, g2p1, m2g);
    } else {
      azimuth_sweep<ORDER2, NUM, false, true>(W, x0, xc, xs, n0, ncf, v0, vc, vs, gx0, gxc, gxs,
                                              kR, kM, kR1, kM1, g2p1, m2g);
    }
    // gbase counted each cosine once; both signs of sin(phi) share it
    const float gbR = 2.0f * W.gbase[0], gbM = 2.0f * W.gbase[1];

    // ---- contraction with the spectral tables (broadcast shared-memory reads) ---------------
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      float tR, tM;
      {
        const float4* rowA = reinterpret_cast<const float4*>(&sA[0][l][c][0]);
        float vA[NUM];
#pragma unroll
        for (int q4 = 0; q4 < NUM / 4; ++q4) {
          const float4 t4 = rowA[q4];
          vA[4 * q4] = t4.x; vA[4 * q4 + 1] = t4.y; vA[4 * q4 + 2] = t4.z; vA[4 * q4 + 3] = t4.w;
        }
        if (!ORDER2) {
          tR = vA[0] * gbR;
          tM = vA[0] * gbM;
#pragma unroll
          for (int s = 0; s < NUM - 1; ++s) {
            tR = fmaf(vA[1 + s], W.ramp[0][s], tR);
            tM = fmaf(vA[1 + s], W.ramp[1][s], tM);
          }
        } else {
          const float4* rowB = reinterpret_cast<const float4*>(&sA[NT - 1][l][c][0]);
          float vB[NUM];
#pragma unroll
          for (int q4 = 0; q4 < NUM / 4; ++q4) {
            const float4 t4 = rowB[q4];
            vB[4 * q4] = t4.x; vB[4 * q4 + 1] = t4.y; vB[4 * q4 + 2] = t4.z; vB[4 * q4 + 3] = t4.w;
          }
          tR = fmaf(vA[0], W.base[0], vB[0] * W.base[2]);
          tM = fmaf(vA[0], W.base[1], vB[0] * W.base[3]);
#pragma unroll
          for (int s = 0; s < NUM - 1; ++s) {
            tR = fmaf(vA[1 + s], W.ramp[0][s], tR);
            tM = fmaf(vA[1 + s], W.ramp[1][s], tM);
            tR = fmaf(vB[1 + s], W.ramp[2][s], tR);
            tM = fmaf(vB[1 + s], W.ramp[3][s], tM);
          }
        }
      }
      if (dc.win_n > 0) {
        const float* e0 = sE0 + c * e_pad + dc.win_i0;
        const float* de = sDE + c * e_pad + dc.win_i0;
        float eR = e0[0] * gbR, eM = e0[0] * gbM;
#pragma unroll
        for (int t = 0; t < kNG; ++t) {
          eR = fmaf(de[t], W.gramp[0][t], eR);
          eM = fmaf(de[t], W.gramp[1][t], eM);
        }
        tR = fmaf(sG[l][c], eR, tR);
        tM = fmaf(sG[l][c], eM, tM);
      }
      acc[c] = fmaf(sCR[c], tR, fmaf(sCM[c], tM, acc[c]));
    }

    // ---- wide ground windows (grazing rays, small planets): extra sweeps, kNG ramps each -----
    for (int w0 = kNG; w0 < dc.win_n; w0 += kNG) {
      W.gbase[0] = W.gbase[1] = 0.f;
#pragma unroll
      for (int t = 0; t < kNG; ++t) W.gramp[0][t] = W.gramp[1][t] = 0.f;
      azimuth_sweep<ORDER2, NUM, true, false>(W, x0, xc, xs, n0, ncf, v0, vc, vs, gx0 - (float)w0,
                                              gxc, gxs, kR, kM, kR1, kM1, g2p1, m2g);
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        const float* de = sDE + c * e_pad + dc.win_i0 + w0;
        float eR = 0.f, eM = 0.f;
#pragma unroll
        for (int t = 0; t < kNG; ++t) {
          // ramps past the last knot read the zero padding of sDE
          eR = fmaf(de[t], W.gramp[0][t], eR);
          eM = fmaf(de[t], W.gramp[1][t], eM);
        }
        acc[c] = fmaf(sCR[c] * sG[l][c], eR, fmaf(sCM[c] * sG[l][c], eM, acc[c]));
      }
    }
  }

  // one interleaved texel per thread: CP contiguous floats, padding channels zero
  float4* out = reinterpret_cast<float4*>(dJ + (layer + (size_t)j * width + i_nu * mu_s_n + i_mu_s) * CP);
#pragma unroll
  for (int q = 0; q < CP / 4; ++q) {
    float v[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) v[e] = (4 * q + e < NC) ? acc[(4 * q + e < NC) ? 4 * q + e : 0] : 0.f;
    out[q] = make_float4(v[0], v[1], v[2], v[3]);
  }
}

template <int NC, bool ORDER2, int NUM>
cudaError_t launch_one(const PasGeometry& g, const PasDensityDir* dirs, const float* G,
                       const float* cR, const float* cM, const float* tabA, const float* tabB,
                       const float* dE, float* dJ, int k_begin, int k_end, cudaStream_t stream) {
  const int threads = 256;
  const int texels = g.sz.mu_n * g.sz.nu_n;
  dim3 grid((texels + threads - 1) / threads, g.sz.mu_s_n, k_end - k_begin);
  const size_t dyn = (size_t)2 * NC * (g.sz.e_w + kNG + 1) * sizeof(float);
  auto kern = density_kernel<NC, ORDER2, NUM>;
  if (dyn > 16 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
    if (e != cudaSuccess) return e;
  }
  kern<<<grid, threads, dyn, stream>>>(g, dirs, G, cR, cM, tabA, tabB, dE, dJ, k_begin);
  return cudaGetLastError();
}

template <int NC>
cudaError_t launch_nc(const PasGeometry& g, const PasDensityDir* dirs, const float* G,
                      const float* cR, const float* cM, const float* dR, const float* dM,
                      const float* dS, const float* dE, int order, float* dJ, int k_begin,
                      int k_end, cudaStream_t stream) {
  const bool wide = g.sz.nu_n > 8;
  if (order == 2) {
    return wide ? launch_one<NC, true, 16>(g, dirs, G, cR, cM, dR, dM, dE, dJ, k_begin, k_end, stream)
                : launch_one<NC, true, 8>(g, dirs, G, cR, cM, dR, dM, dE, dJ, k_begin, k_end, stream);
  }
  return wide ? launch_one<NC, false, 16>(g, dirs, G, cR, cM, dS, nullptr, dE, dJ, k_begin, k_end, stream)
              : launch_one<NC, false, 8>(g, dirs, G, cR, cM, dS, nullptr, dE, dJ, k_begin, k_end, stream);
}

}  // namespace

cudaError_t launch_scattering_density(const PasGeometry& g, const PasSpectrum& s,
                                      const PasDensityDir* dirs, const float* G, const float* cR,
                                      const float* cM, const float* dR, const float* dM,
                                      const float* dS, const float* dE, int order, float* dJ,
                                      int k_begin, int k_end, cudaStream_t stream) {
  if (g.sz.nu_n < 2 || g.sz.nu_n > PAS_MAX_NU) return cudaErrorInvalidValue;
  switch (s.nc) {
#define PAS_CASE(N) \
  case N: return launch_nc<N>(g, dirs, G, cR, cM, dR, dM, dS, dE, order, dJ, k_begin, k_end, stream);
    PAS_CASE(3) PAS_CASE(4) PAS_CASE(8) PAS_CASE(15) PAS_CASE(16)
#undef PAS_CASE
    default: return cudaErrorInvalidValue;
  }
}

bool channel_count_supported(int nc) {
  return nc == 3 || nc == 4 || nc == 8 || nc == 15 || nc == 16;
}

}  // namespace pas

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
//
// Ramp window: omega_s = (sx, sy, mu_s) is a unit vector, so over the azimuths of one polar
// direction x stays inside x0 +- A with x0 = (mu_s cos(theta) + 1)(NU - 1)/2 and
// A = sin(theta) sqrt(1 - mu_s^2)(NU - 1)/2 -- the same interval for every texel of the block
// (mu_s is the block's column). Ramps below the interval are identically 1 and telescope into the
// base value (V[0] + sum_{s < lo} D[s] = V[lo]); ramps above it are identically 0. The rows are
// therefore staged REBASED at lo = floor(x0 - A): entry 0 = V[lo], entry 1 + s = D[lo + s], and a
// direction sweeps only the entry pairs its window needs (5.0 of the 7 ramps on average at the
// reference's sizes; 1..4 register pairs instead of always 4).
#ifndef PAS_B200_CSRC_KERNEL_DENSITY_CUH_
#define PAS_B200_CSRC_KERNEL_DENSITY_CUH_

#include <cooperative_groups.h>

#include "pas_kernels.h"
#include "pas_physics.cuh"

namespace pas {
namespace density {

constexpr int kNG = 4;  // irradiance ramps handled per sweep of the azimuth loop
// measured on B200 (15 channels): orders >= 3 0.714 ms at 4, 0.729 at 2 / 8, 0.80 fully unrolled; order 2
// (twice the code per cosine) 1.318 ms at 2, 1.347 at 4, 1.49 at 8, 1.76 fully unrolled
#ifndef PAS_SWEEP_UNROLL
#define PAS_SWEEP_UNROLL 4
#endif
#ifndef PAS_SWEEP_UNROLL_ORDER2
#define PAS_SWEEP_UNROLL_ORDER2 2
#endif

// cos/sin of phi_m = (m + 0.5) pi / 16, m = 0..15 (functions.glsl:1216); m' = 31 - m mirrors sin.
__device__ constexpr float kCosPhi[16] = {
    0.99518472f, 0.956940353f, 0.881921291f, 0.773010433f, 0.634393275f, 0.471396744f,
    0.290284663f, 0.0980171412f, -0.0980171412f, -0.290284663f, -0.471396744f, -0.634393275f,
    -0.773010433f, -0.881921291f, -0.956940353f, -0.99518472f};
__device__ constexpr float kSinPhi[16] = {
    0.0980171412f, 0.290284663f, 0.471396744f, 0.634393275f, 0.773010433f, 0.881921291f,
    0.956940353f, 0.99518472f, 0.99518472f, 0.956940353f, 0.881921291f, 0.773010433f,
    0.634393275f, 0.471396744f, 0.290284663f, 0.0980171412f};

// The same table in constant memory for the partially unrolled sweep: its loop counter is warp uniform,
// so the pair is fetched through the uniform datapath.
static __constant__ float2 kPhi[16] = {
    {0.99518472f, 0.0980171412f}, {0.956940353f, 0.290284663f}, {0.881921291f, 0.471396744f},
    {0.773010433f, 0.634393275f}, {0.634393275f, 0.773010433f}, {0.471396744f, 0.881921291f},
    {0.290284663f, 0.956940353f}, {0.0980171412f, 0.99518472f}, {-0.0980171412f, 0.99518472f},
    {-0.290284663f, 0.956940353f}, {-0.471396744f, 0.881921291f}, {-0.634393275f, 0.773010433f},
    {-0.773010433f, 0.634393275f}, {-0.881921291f, 0.471396744f}, {-0.956940353f, 0.290284663f},
    {-0.99518472f, 0.0980171412f}};

struct DirConst {       // per (block, l), in shared memory
  float cos_t, sin_t;
  float g_a, g_b;       // ground cosine -> irradiance texel x: x = g_a + g_b * nu1
  int hit;
  int win_i0;           // first irradiance knot of the window
  int win_n;            // number of ramps needed (>= 0)
  int nu_lo;            // first nu knot of the ramp window; packed with the entry pairs to sweep:
                        //   nu_lo | (pairs << 16)
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

// ---- packed-fp32 formulation ---------------------------------------------------------------------
// Same algorithm; the kernel above is bound by instruction issue (ncu: issue slots 79 % busy, FMA
// pipe 65 %), so this one halves the number of FMA instructions with Blackwell's packed FFMA2
// (fma.rn.f32x2, two independent fp32 FMAs per instruction, results identical to two FFMA):
//  * sweep: the NUM - 1 ramps and the base term form NUM "entries", accumulated as NUM / 2 register
//    pairs (entry NUM - 1 is the base: its "ramp" is the constant 1);
//  * contraction: channels are processed in pairs; the tables are staged in shared memory with the
//    two channels of a pair side by side, [l][pair][knot] float2, and the weights are broadcast to
//    both halves;
//  * the ground window is swept with 0, 2 or 4 ramps, whichever covers it (most directions that
//    reach the ground see less than two irradiance texels).
__device__ __forceinline__ float2 dup2(float v) { return make_float2(v, v); }
template <bool ORDER2, int NE, int NG>
struct Weights2 {
  // Entry i of a weight set pairs with element i of a staged row: i = 0 is the base (knot-0 value,
  // "ramp" = 1), i >= 1 the ramp of knot difference i - 1. Entries 2h and 2h + 1 share a register pair.
  // Order >= 3: set 0 pairs with the Rayleigh coefficient, set 1 with the Mie one.
  // Order 2: sets 0..3 = (table R, coef R), (table R, coef M), (table M, coef R), (table M, coef M).
  static constexpr int NW = ORDER2 ? 4 : 2;
  float2 w[NW][NE];                   // NE entry pairs = the base + up to 2 NE - 1 ramps of the window
  float2 gb;                          // ORDER2: (sum pR, sum pM), base of the ground term
  float2 g[2][NG > 0 ? NG / 2 : 1];   // ground ramps, coef R / coef M
};

template <bool ORDER2, int NE, int NG>
__device__ __forceinline__ void azimuth_sweep2(Weights2<ORDER2, NE, NG>& W, float x0, float xc,
                                               float xs, float n0, float nc_, float v0, float vc,
                                               float vs, float gx0, float gxc, float gxs, float kR,
                                               float kM, float kR1, float kM1, float g2p1, float m2g) {
  // Partially unrolled: the block runs up to a dozen (entry pairs, ground ramps) variants of this sweep,
  // one per polar direction; fully unrolled they do not fit the instruction cache together (measured:
  // hit rate 97 % -> 73 %, "no instruction" the top stall). Taking the directions of a block in the order
  // of their variants instead was measured and does not help.
  constexpr int kSweepUnroll = ORDER2 ? PAS_SWEEP_UNROLL_ORDER2 : PAS_SWEEP_UNROLL;
#pragma unroll kSweepUnroll
  for (int m = 0; m < 16; ++m) {
    const float2 phi = kPhi[m];
    const float c = phi.x, s = phi.y;
    const float nu2 = fmaf(nc_, c, n0);
    const float q = fmaf(nu2, nu2, 1.0f);
    const float pR = kR * q;
    const float rs = f_rsqrt(fmaf(m2g, nu2, g2p1));
    const float pM = (kM * q) * (rs * rs) * rs;
    const float2 pR2 = dup2(pR), pM2 = dup2(pM);
    const float xb = fmaf(xc, c, x0);
    const float xp = fmaf(xs, s, xb), xm = fmaf(-xs, s, xb);
    if (!ORDER2) {
#pragma unroll
      for (int h = 0; h < NE; ++h) {
        // entries 2h (ramp 2h - 1, or the base: counted once per cosine, doubled by the caller) and
        // 2h + 1 (ramp 2h); x is relative to the first knot of the window
        float2 cs;
        cs.x = h == 0 ? 1.0f : f_sat(xp - (float)(2 * h - 1)) + f_sat(xm - (float)(2 * h - 1));
        cs.y = f_sat(xp - (float)(2 * h)) + f_sat(xm - (float)(2 * h));
        W.w[0][h] = __ffma2_rn(cs, pR2, W.w[0][h]);
        W.w[1][h] = __ffma2_rn(cs, pM2, W.w[1][h]);
      }
    } else {
      // order 2: incident radiance = R * P_R(nu1) + M * P_M(nu1) (functions.glsl:995-1003)
      const float vb = fmaf(vc, c, v0);
      const float nu1p = fmaf(vs, s, vb), nu1m = fmaf(-vs, s, vb);
      const float q1p = fmaf(nu1p, nu1p, 1.0f), q1m = fmaf(nu1m, nu1m, 1.0f);
      const float rp = f_rsqrt(fmaf(m2g, nu1p, g2p1)), rm = f_rsqrt(fmaf(m2g, nu1m, g2p1));
      const float2 PRp2 = dup2(kR1 * q1p), PRm2 = dup2(kR1 * q1m);
      const float2 PMp2 = dup2((kM1 * q1p) * (rp * rp) * rp), PMm2 = dup2((kM1 * q1m) * (rm * rm) * rm);
#pragma unroll
      for (int h = 0; h < NE; ++h) {
        float2 cp, cm;
        cp.x = h == 0 ? 1.0f : f_sat(xp - (float)(2 * h - 1));
        cm.x = h == 0 ? 1.0f : f_sat(xm - (float)(2 * h - 1));
        cp.y = f_sat(xp - (float)(2 * h));
        cm.y = f_sat(xm - (float)(2 * h));
        const float2 a = __ffma2_rn(cp, PRp2, __fmul2_rn(cm, PRm2));  // table R weight
        const float2 b = __ffma2_rn(cp, PMp2, __fmul2_rn(cm, PMm2));  // table M weight
        W.w[0][h] = __ffma2_rn(a, pR2, W.w[0][h]);
        W.w[1][h] = __ffma2_rn(a, pM2, W.w[1][h]);
        W.w[2][h] = __ffma2_rn(b, pR2, W.w[2][h]);
        W.w[3][h] = __ffma2_rn(b, pM2, W.w[3][h]);
      }
      if (NG > 0) W.gb = __fadd2_rn(W.gb, make_float2(pR, pM));
    }
    if (NG > 0) {
      const float gb = fmaf(gxc, c, gx0);
      const float gp = fmaf(gxs, s, gb), gm = fmaf(-gxs, s, gb);
#pragma unroll
      for (int t = 0; t < NG; t += 2) {
        float2 cs;
        cs.x = f_sat(gp - (float)t) + f_sat(gm - (float)t);
        cs.y = f_sat(gp - (float)(t + 1)) + f_sat(gm - (float)(t + 1));
        W.g[0][t / 2] = __ffma2_rn(cs, pR2, W.g[0][t / 2]);
        W.g[1][t / 2] = __ffma2_rn(cs, pM2, W.g[1][t / 2]);
      }
    }
  }
}

__device__ __forceinline__ float half_of(const float2& v, int i) { return i ? v.y : v.x; }

// Contraction of the entries [CH0, min(CH0 + 8, 2 NE)) of weight sets (2 TAB, 2 TAB + 1) with the staged
// rows of table TAB (NUM entries per channel pair), for every channel pair; the pass that holds entry 0
// of table 0 also adds the ground term.
template <int NP, bool ORDER2, int NUM, int NE, int NG, int TAB, int CH0>
__device__ __forceinline__ void contract_pass(float2 (&acc)[NP], const Weights2<ORDER2, NE, NG>& W,
                                              const float2* __restrict__ row,
                                              const float2* __restrict__ sG_l,
                                              const float2* __restrict__ sCR,
                                              const float2* __restrict__ sCM,
                                              const float2* __restrict__ e0p,
                                              const float2* __restrict__ dep, int e_pad) {
  constexpr bool GROUND = NG > 0 && TAB == 0 && CH0 == 0;
  constexpr int N = 2 * NE - CH0 < 8 ? 2 * NE - CH0 : 8;   // entries of this pass (even)
  float2 dR[N], dM[N];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const int e = CH0 + i;
    float r = half_of(W.w[2 * TAB][e / 2], e & 1), m = half_of(W.w[2 * TAB + 1][e / 2], e & 1);
    // order >= 3: the base entry counted each cosine once; both signs of sin(phi) share it
    if (!ORDER2 && e == 0) { r *= 2.0f; m *= 2.0f; }
    dR[i] = dup2(r);
    dM[i] = dup2(m);
  }
  float2 gbR2 = dR[0], gbM2 = dM[0];
  float2 dgR[NG > 0 ? NG : 1], dgM[NG > 0 ? NG : 1];
  if (GROUND) {
    if (ORDER2) {
      gbR2 = dup2(2.0f * W.gb.x);
      gbM2 = dup2(2.0f * W.gb.y);
    }
#pragma unroll
    for (int t = 0; t < NG; ++t) {
      dgR[t] = dup2(half_of(W.g[0][t / 2], t & 1));
      dgM[t] = dup2(half_of(W.g[1][t / 2], t & 1));
    }
  }
#pragma unroll
  for (int cp = 0; cp < NP; ++cp) {
    float2 v[N];
    const float4* r4 = reinterpret_cast<const float4*>(row + cp * NUM + CH0);
#pragma unroll
    for (int h = 0; h < N / 2; ++h) {
      const float4 t4 = r4[h];
      v[2 * h] = make_float2(t4.x, t4.y);
      v[2 * h + 1] = make_float2(t4.z, t4.w);
    }
    float2 tR = __fmul2_rn(v[0], dR[0]), tM = __fmul2_rn(v[0], dM[0]);
#pragma unroll
    for (int i = 1; i < N; ++i) {
      tR = __ffma2_rn(v[i], dR[i], tR);
      tM = __ffma2_rn(v[i], dM[i], tM);
    }
    if (GROUND) {
      const float2* e0 = e0p + cp * e_pad;
      const float2* de = dep + cp * e_pad;
      float2 eR = __fmul2_rn(e0[0], gbR2), eM = __fmul2_rn(e0[0], gbM2);
#pragma unroll
      for (int t = 0; t < NG; ++t) {
        const float2 dv = de[t];
        eR = __ffma2_rn(dv, dgR[t], eR);
        eM = __ffma2_rn(dv, dgM[t], eM);
      }
      const float2 G2 = sG_l[cp];
      tR = __ffma2_rn(G2, eR, tR);
      tM = __ffma2_rn(G2, eM, tM);
    }
    acc[cp] = __ffma2_rn(sCR[cp], tR, __ffma2_rn(sCM[cp], tM, acc[cp]));
  }
}

// One polar direction: sweep + contraction into acc[pair].
template <int NP, bool ORDER2, int NUM, int NE, int NG>
__device__ __forceinline__ void direction2(float2 (&acc)[NP], const float2* __restrict__ rowA,
                                           const float2* __restrict__ rowB,
                                           const float2* __restrict__ sG_l,
                                           const float2* __restrict__ sCR,
                                           const float2* __restrict__ sCM,
                                           const float2* __restrict__ e0p,
                                           const float2* __restrict__ dep, int e_pad, float x0,
                                           float xc, float xs, float n0, float ncf, float v0, float vc,
                                           float vs, float gx0, float gxc, float gxs, float kR,
                                           float kM, float kR1, float kM1, float g2p1, float m2g) {
  static_assert((NUM == 8 || NUM == 16) && NE >= 1 && 2 * NE <= NUM, "entries are contracted in chunks of 8");
  Weights2<ORDER2, NE, NG> W;
#pragma unroll
  for (int a = 0; a < Weights2<ORDER2, NE, NG>::NW; ++a) {
#pragma unroll
    for (int e = 0; e < NE; ++e) W.w[a][e] = make_float2(0.f, 0.f);
  }
  W.gb = make_float2(0.f, 0.f);
#pragma unroll
  for (int t = 0; t < (NG > 0 ? NG / 2 : 1); ++t) W.g[0][t] = W.g[1][t] = make_float2(0.f, 0.f);
  azimuth_sweep2<ORDER2, NE, NG>(W, x0, xc, xs, n0, ncf, v0, vc, vs, gx0, gxc, gxs, kR, kM, kR1, kM1,
                                 g2p1, m2g);
  contract_pass<NP, ORDER2, NUM, NE, NG, 0, 0>(acc, W, rowA, sG_l, sCR, sCM, e0p, dep, e_pad);
  if (2 * NE > 8) contract_pass<NP, ORDER2, NUM, NE, NG, 0, (2 * NE > 8 ? 8 : 0)>(acc, W, rowA, sG_l, sCR, sCM, e0p, dep, e_pad);
  if (ORDER2) {
    contract_pass<NP, ORDER2, NUM, NE, NG, ORDER2 ? 1 : 0, 0>(acc, W, rowB, sG_l, sCR, sCM, e0p, dep, e_pad);
    if (2 * NE > 8) {
      contract_pass<NP, ORDER2, NUM, NE, NG, ORDER2 ? 1 : 0, (2 * NE > 8 ? 8 : 0)>(acc, W, rowB, sG_l, sCR, sCM, e0p, dep, e_pad);
    }
  }
}

// One polar direction with the ground window swept with 0, 2 or 4 ramps, whichever covers it.
template <int NP, bool ORDER2, int NUM, int NE>
__device__ __forceinline__ void direction_ng(int win_n, float2 (&acc)[NP], const float2* __restrict__ rowA,
                                             const float2* __restrict__ rowB, const float2* __restrict__ sG_l,
                                             const float2* __restrict__ sCR, const float2* __restrict__ sCM,
                                             const float2* __restrict__ e0p, const float2* __restrict__ dep,
                                             int e_pad, float x0, float xc, float xs, float n0, float ncf,
                                             float v0, float vc, float vs, float gx0, float gxc, float gxs,
                                             float kR, float kM, float kR1, float kM1, float g2p1, float m2g) {
  if (win_n == 0) {
    direction2<NP, ORDER2, NUM, NE, 0>(acc, rowA, rowB, sG_l, sCR, sCM, e0p, dep, e_pad, x0, xc, xs, n0, ncf,
                                       v0, vc, vs, gx0, gxc, gxs, kR, kM, kR1, kM1, g2p1, m2g);
  } else if (win_n <= 2) {
    direction2<NP, ORDER2, NUM, NE, 2>(acc, rowA, rowB, sG_l, sCR, sCM, e0p, dep, e_pad, x0, xc, xs, n0, ncf,
                                       v0, vc, vs, gx0, gxc, gxs, kR, kM, kR1, kM1, g2p1, m2g);
  } else {
    direction2<NP, ORDER2, NUM, NE, 4>(acc, rowA, rowB, sG_l, sCR, sCM, e0p, dep, e_pad, x0, xc, xs, n0, ncf,
                                       v0, vc, vs, gx0, gxc, gxs, kR, kM, kR1, kM1, g2p1, m2g);
  }
}

// CP: channel pitch of the tables (4, 8 or 16 floats per texel; nc <= CP channels are live, the rest
// stay zero). MIRROR: multi-GPU instantiation with the transposed output path.
template <int CP, bool ORDER2, int NUM, bool MIRROR>
__global__ void __launch_bounds__(256, 2)
density_kernel_x2(const __grid_constant__ PasGeometry g, int NC, const PasDensityDir* __restrict__ dirs,
                  const float* __restrict__ G, const float* __restrict__ cRk,
                  const float* __restrict__ cMk, const float* __restrict__ tabA,
                  const float* __restrict__ tabB, const float* __restrict__ dE,
                  float* __restrict__ dJ, const __grid_constant__ PeerTables mirrors, int k_begin,
                  int k_stride) {
  constexpr int NT = ORDER2 ? 2 : 1;
  constexpr int NP = CP / 2;
  extern __shared__ __align__(16) float smem_dyn[];
  __shared__ __align__(16) float2 sA[NT][PAS_DIR_THETA][NP][NUM];
  __shared__ float2 sG[PAS_DIR_THETA][NP];
  __shared__ float2 sCR[NP], sCM[NP];
  __shared__ DirConst sDir[PAS_DIR_THETA];

  const int tid = threadIdx.x;
  const int i_mu_s = blockIdx.y;
  const int k = k_begin + blockIdx.z * k_stride;
  const int mu_n = g.sz.mu_n, nu_n = g.sz.nu_n, mu_s_n = g.sz.mu_s_n, e_w = g.sz.e_w;
  const int width = nu_n * mu_s_n;
  const size_t layer = (size_t)k * mu_n * width;
  // dynamic shared memory: irradiance row 0 and its forward differences, zero padded, channel pairs
  const int e_pad = e_w + kNG + 1;
  float2* sE0 = reinterpret_cast<float2*>(smem_dyn);  // [NP][e_pad]
  float2* sDE = sE0 + NP * e_pad;                     // [NP][e_pad]

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
    const double half = 0.5 * (e_w - 1);
    const double ga = (r * mu_s_d / g.bottom) * half + half;
    const double gb = (double)d.dg_over_b * half;
    dc.g_a = (float)ga;
    dc.g_b = (float)gb;
    int i0 = (int)floor(ga - gb - 1e-3);
    i0 = i0 < 0 ? 0 : (i0 > e_w - 2 ? e_w - 2 : i0);
    int i1 = (int)ceil(ga + gb + 1e-3);
    i1 = i1 > e_w - 1 ? e_w - 1 : (i1 < i0 + 1 ? i0 + 1 : i1);
    dc.win_i0 = i0;
    dc.win_n = d.hit ? i1 - i0 : 0;
    // nu ramp window of (column, direction): x in x0 +- A for every texel and azimuth (header comment)
    {
      const double sc = 0.5 * (nu_n - 1);
      const double xw0 = (mu_s_d * (double)d.cos_t + 1.0) * sc;
      const double aw = sc * (double)d.sin_t * sqrt(d_pos(1.0 - mu_s_d * mu_s_d)) + 1e-3;
      int lo = (int)floor(xw0 - aw);
      lo = lo < 0 ? 0 : (lo > nu_n - 2 ? nu_n - 2 : lo);
      int hi = (int)ceil(xw0 + aw);                      // ramps lo .. hi - 1 are live
      hi = hi > nu_n - 1 ? nu_n - 1 : (hi < lo + 1 ? lo + 1 : hi);
      int pairs = (hi - lo + 2) / 2;                     // base + (hi - lo) ramps, two entries per pair
      if (NUM == 16) pairs = (pairs + 1) & ~1;           // wide tables: 2, 4, 6 or 8 pairs are instantiated
      dc.nu_lo = lo | (pairs << 16);
    }
    sDir[tid] = dc;
  }
  {
    float* fCR = reinterpret_cast<float*>(sCR);
    float* fCM = reinterpret_cast<float*>(sCM);
    if (tid < CP) {
      fCR[tid] = tid < NC ? cRk[k * PAS_MAX_CH + tid] : 0.f;
      fCM[tid] = tid < NC ? cMk[k * PAS_MAX_CH + tid] : 0.f;
    }
    float* fG = reinterpret_cast<float*>(sG);
    for (int idx = tid; idx < PAS_DIR_THETA * CP; idx += blockDim.x) {
      const int l = idx / CP, c = idx % CP;
      fG[idx] = c < NC ? G[(size_t)(k * PAS_DIR_THETA + l) * PAS_MAX_CH + c] : 0.f;
    }
    float* fE0 = reinterpret_cast<float*>(sE0);
    float* fDE = reinterpret_cast<float*>(sDE);
    for (int idx = tid; idx < CP * e_pad; idx += blockDim.x) {
      const int c = idx / e_pad, i = idx % e_pad;
      const float* row = dE + (size_t)c * e_w * g.sz.e_h;  // row 0: r = bottom
      const int o = ((c >> 1) * e_pad + i) * 2 + (c & 1);
      fE0[o] = (c < NC && i < e_w) ? row[i] : 0.f;
      fDE[o] = (c < NC && i < e_w - 1) ? row[i + 1] - row[i] : 0.f;
    }
    // mu-interpolated rows: value at slab s, then in place -> (V[0], D[0..NUM-2]). Consecutive
    // threads read the consecutive channels of one interleaved texel (64 B at 15 channels).
    float* fA = reinterpret_cast<float*>(sA);
    for (int idx = tid; idx < NT * PAS_DIR_THETA * NUM * CP; idx += blockDim.x) {
      const int c = idx % CP, s = (idx / CP) % NUM, l = (idx / (CP * NUM)) % PAS_DIR_THETA;
      const int t = idx / (CP * NUM * PAS_DIR_THETA);
      float v = 0.f;
      if (s < nu_n && c < NC) {
        const PasDensityDir d = dirs[k * PAS_DIR_THETA + l];
        const float* tab = (t == 0 ? tabA : tabB) + (layer + s * mu_s_n + i_mu_s) * CP + c;
        const float a = tab[(size_t)d.j0 * width * CP], b = tab[(size_t)d.j1 * width * CP];
        v = fmaf(d.w_row, b - a, a);
      }
      fA[((((t * PAS_DIR_THETA + l) * NP + (c >> 1)) * NUM) + s) * 2 + (c & 1)] = v;
    }
  }
  __syncthreads();
  // in place: knot values -> rebased entries (V[lo], D[lo], D[lo + 1], ..., 0, ...) of the direction's window
  for (int row = tid; row < NT * PAS_DIR_THETA * NP; row += blockDim.x) {
    float2* p = (&sA[0][0][0][0]) + row * NUM;
    const int lo = sDir[(row / NP) % PAS_DIR_THETA].nu_lo & 0xffff;
    float2 v[NUM];
#pragma unroll
    for (int s = 0; s < NUM; ++s) v[s] = p[s];
    float2 base = v[0];
#pragma unroll
    for (int s = 1; s < NUM; ++s) {
      if (s == lo) base = v[s];
    }
    p[0] = base;
#pragma unroll
    for (int s = 1; s < NUM; ++s) p[s] = make_float2(0.f, 0.f);
#pragma unroll
    for (int s = 0; s < NUM - 1; ++s) {
      // D[s] goes to entry 1 + s - lo; differences below the window have telescoped into the base
      if (s >= lo && s + 1 < nu_n) p[1 + s - lo] = make_float2(v[s + 1].x - v[s].x, v[s + 1].y - v[s].y);
    }
  }
  __syncthreads();

  // ---- per-texel geometry (fp64, once) --------------------------------------------------------
  constexpr int Q = CP / 4;
  const int texel = blockIdx.x * blockDim.x + tid;
  const bool active = texel < mu_n * nu_n;
  // (the multi-GPU instantiation ends with cluster-wide barriers: its idle threads tag along on texel 0)
  if (!active && !(MIRROR && Q > 1)) return;
  const int j = active ? texel / nu_n : 0, i_nu = active ? texel % nu_n : 0;
  double mu_d, r_mu_d;
  bool hit_unused;
  scattering_row_mu(g, r, rho, j, &mu_d, &r_mu_d, &hit_unused);
  const double nu_d = scattering_slab_nu(g, i_nu, mu_d, mu_s_d);
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

  float2 acc[NP];
#pragma unroll
  for (int c = 0; c < NP; ++c) acc[c] = make_float2(0.f, 0.f);

#pragma unroll 1
  for (int l = 0; l < PAS_DIR_THETA; ++l) {
    const DirConst dc = sDir[l];
    const float ct = dc.cos_t, st = dc.sin_t;
    const float domega = dtheta_dphi * st;  // functions.glsl:1219
    const float v0 = mu_s * ct, vc = sx * st, vs = sy * st;
    const int nu_lo = dc.nu_lo & 0xffff, pairs = dc.nu_lo >> 16;
    const float x0 = fmaf(v0, scale, scale) - (float)nu_lo, xc = vc * scale, xs = vs * scale;
    const float n0 = mu * ct, ncf = wx * st;
    const float kR = kR1 * domega, kM = kM1 * domega;
    const float gx0 = fmaf(dc.g_b, v0, dc.g_a) - (float)dc.win_i0;
    const float gxc = dc.g_b * vc, gxs = dc.g_b * vs;
    const float2* rowA = &sA[0][l][0][0];
    const float2* rowB = &sA[NT - 1][l][0][0];
    const float2* e0p = sE0 + dc.win_i0;
    const float2* dep = sDE + dc.win_i0;
#define PAS_DIRECTION(NE_)                                                                                   \
    direction_ng<NP, ORDER2, NUM, NE_>(dc.win_n, acc, rowA, rowB, sG[l], sCR, sCM, e0p, dep, e_pad, x0, xc,  \
                                       xs, n0, ncf, v0, vc, vs, gx0, gxc, gxs, kR, kM, kR1, kM1, g2p1, m2g)
    if (NUM == 8) {
      switch (pairs) {
        case 1: PAS_DIRECTION(1); break;
        case 2: PAS_DIRECTION(2); break;
        case 3: PAS_DIRECTION(3); break;
        default: PAS_DIRECTION(4); break;
      }
    } else {
      switch (pairs) {
        case 2: PAS_DIRECTION(2); break;
        case 4: PAS_DIRECTION(4); break;
        case 6: PAS_DIRECTION(NUM == 16 ? 6 : 1); break;
        default: PAS_DIRECTION(NUM == 16 ? 8 : 1); break;
      }
    }
#undef PAS_DIRECTION
    // ---- wide ground windows (grazing rays, small planets): extra sweeps, kNG ramps each -----
    for (int w0 = kNG; w0 < dc.win_n; w0 += kNG) {
      Weights<ORDER2, NUM> W;
      W.gbase[0] = W.gbase[1] = 0.f;
#pragma unroll
      for (int t = 0; t < kNG; ++t) W.gramp[0][t] = W.gramp[1][t] = 0.f;
      azimuth_sweep<ORDER2, NUM, true, false>(W, x0, xc, xs, n0, ncf, v0, vc, vs, gx0 - (float)w0,
                                              gxc, gxs, kR, kM, kR1, kM1, g2p1, m2g);
      float2 dgR[kNG], dgM[kNG];
#pragma unroll
      for (int t = 0; t < kNG; ++t) {
        dgR[t] = dup2(W.gramp[0][t]);
        dgM[t] = dup2(W.gramp[1][t]);
      }
#pragma unroll
      for (int cp = 0; cp < NP; ++cp) {
        const float2* de = dep + cp * e_pad + w0;
        float2 eR = make_float2(0.f, 0.f), eM = make_float2(0.f, 0.f);
#pragma unroll
        for (int t = 0; t < kNG; ++t) {
          // ramps past the last knot read the zero padding of sDE
          const float2 dv = de[t];
          eR = __ffma2_rn(dv, dgR[t], eR);
          eM = __ffma2_rn(dv, dgM[t], eM);
        }
        const float2 G2 = sG[l][cp];
        acc[cp] = __ffma2_rn(__fmul2_rn(sCR[cp], G2), eR, __ffma2_rn(__fmul2_rn(sCM[cp], G2), eM, acc[cp]));
      }
    }
  }

  // ---- output: one interleaved texel (CP floats) per thread ---------------------------------------
  if (!MIRROR || Q == 1) {
    const size_t offset = (layer + (size_t)j * width + i_nu * mu_s_n + i_mu_s) * CP;
#pragma unroll
    for (int q = 0; q < Q; ++q) {
      const float4 v = make_float4(acc[2 * q].x, acc[2 * q].y, acc[2 * q + 1].x, acc[2 * q + 1].y);
      reinterpret_cast<float4*>(dJ + offset)[q] = v;
      if (MIRROR) {
        for (int p = 0; p < mirrors.n; ++p) reinterpret_cast<float4*>(mirrors.tab[p] + offset)[q] = v;
      }
    }
    return;
  }
  // Multi-GPU: every rank needs this layer for its multiple-scattering rays, so the
  // texels are also stored into the tables of the other GPUs: posted stores over NVLink, completed by
  // the barrier kernel that follows the pass. The texels of one block are 2 KB apart (same mu_s
  // column, consecutive (mu, nu)); stored thread by thread they would cross the link as isolated
  // 16-byte packets. The CL blocks of a thread-block cluster own CL neighbouring columns, i.e.
  // CL x 64 contiguous bytes per (mu, nu): each block parks its texels in shared memory, and after a
  // cluster barrier each block writes 1 / CL of the (mu, nu) range for ALL the columns of the
  // cluster, reading the other blocks' texels through distributed shared memory. CL * Q
  // neighbouring lanes then store one contiguous run of CL * 64 bytes (256 B at CL = 4).
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  const int CL = (int)cluster.dim_blocks().y;
  const int crank = (int)cluster.block_rank();       // cluster dims are (1, CL, 1)
  float4* stage = reinterpret_cast<float4*>(sDE + NP * e_pad);  // [256][Q + 1]
#pragma unroll
  for (int q = 0; q < Q; ++q) {
    stage[tid * (Q + 1) + q] = make_float4(acc[2 * q].x, acc[2 * q].y, acc[2 * q + 1].x, acc[2 * q + 1].y);
  }
  cluster.sync();
  const int per = (int)blockDim.x / CL;              // (mu, nu) texels written by this block
  const int col0 = i_mu_s - crank;                   // first column of the cluster
  for (int idx = tid; idx < per * CL * Q; idx += blockDim.x) {
    const int q = idx % Q, c = (idx / Q) % CL, t = crank * per + idx / (Q * CL);
    const int tex = blockIdx.x * blockDim.x + t;
    if (tex >= mu_n * nu_n) continue;
    const float4* src = cluster.map_shared_rank(stage, c);
    const float4 v = src[t * (Q + 1) + q];
    const size_t offset = (layer + (size_t)(tex / nu_n) * width + (tex % nu_n) * mu_s_n + col0 + c) * CP + 4 * q;
    *reinterpret_cast<float4*>(dJ + offset) = v;
    for (int p = 0; p < mirrors.n; ++p) *reinterpret_cast<float4*>(mirrors.tab[p] + offset) = v;
  }
  cluster.sync();  // the shared memory of a block must outlive the reads of its cluster mates
}

template <int CP, bool ORDER2, int NUM, bool MIRROR>
cudaError_t launch_one(const PasGeometry& g, int nc, const PasDensityDir* dirs, const float* G,
                       const float* cR, const float* cM, const float* tabA, const float* tabB,
                       const float* dE, float* dJ, const PeerTables& mirrors, LayerSet layers,
                       cudaStream_t stream) {
  const int texels = g.sz.mu_n * g.sz.nu_n;
  if (layers.count() == 0) return cudaSuccess;
  // (half-size blocks for the small grids of an 8-GPU world -- 1024 blocks of 128 threads instead of 512
  // of 256 on 148 SMs -- were measured: no difference, 0.1356 vs 0.1345 ms)
  const int threads = 256;
  dim3 grid((texels + threads - 1) / threads, g.sz.mu_s_n, layers.count());
  constexpr int Q = CP / 4;
  // irradiance row + differences, then (MIRROR) the output staging [256][Q + 1] float4
  const size_t dyn = (size_t)2 * CP * (g.sz.e_w + kNG + 1) * sizeof(float) +
                     (MIRROR && Q > 1 ? (size_t)(threads / 32) * 32 * (Q + 1) * sizeof(float4) : 0);
  auto kern = density_kernel_x2<CP, ORDER2, NUM, MIRROR>;
  if (dyn > 16 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
    if (e != cudaSuccess) return e;
  }
  if (MIRROR && Q > 1) {
    // clusters of 2 (or 1) neighbouring mu_s columns: see the output stage of the kernel
    // measured on 8 / 4 / 2 B200 (15 channels): pairs of columns (128-byte runs) beat single columns
    // from 3 mirrors on (1.54 vs 1.69 ms, 2.31 vs 2.35 ms); clusters of 4 lose to the longer wait at
    // the cluster barrier (1.59 ms); with one mirror the plain layout wins (3.96 vs 4.08 ms)
    const unsigned cl = ((mirrors.n >= 3 || mirrors.multicast) && g.sz.mu_s_n % 2 == 0) ? 2 : 1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = dyn;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 1;
    attr[0].val.clusterDim.y = cl;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, g, nc, dirs, G, cR, cM, tabA, tabB, dE, dJ, mirrors, layers.begin,
                              layers.stride);
  }
  kern<<<grid, threads, dyn, stream>>>(g, nc, dirs, G, cR, cM, tabA, tabB, dE, dJ, mirrors, layers.begin,
                              layers.stride);
  return cudaGetLastError();
}

template <int CP, bool MIRROR>
cudaError_t launch_cp(const PasGeometry& g, int nc, const PasDensityDir* dirs, const float* G,
                      const float* cR, const float* cM, const float* dR, const float* dM,
                      const float* dS, const float* dE, int order, float* dJ,
                      const PeerTables& mirrors, LayerSet layers, cudaStream_t stream) {
  const bool wide = g.sz.nu_n > 8;
  if (order == 2) {
    return wide ? launch_one<CP, true, 16, MIRROR>(g, nc, dirs, G, cR, cM, dR, dM, dE, dJ, mirrors, layers, stream)
                : launch_one<CP, true, 8, MIRROR>(g, nc, dirs, G, cR, cM, dR, dM, dE, dJ, mirrors, layers, stream);
  }
  return wide ? launch_one<CP, false, 16, MIRROR>(g, nc, dirs, G, cR, cM, dS, nullptr, dE, dJ, mirrors, layers, stream)
              : launch_one<CP, false, 8, MIRROR>(g, nc, dirs, G, cR, cM, dS, nullptr, dE, dJ, mirrors, layers, stream);
}

}  // namespace density
}  // namespace pas

#endif  // PAS_B200_CSRC_KERNEL_DENSITY_CUH_

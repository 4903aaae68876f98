// The two ray-march passes: single scattering (ComputeSingleScatteringTexture,
// atmosphere/functions.glsl:650-730, 933-945) and multiple scattering
// (ComputeMultipleScatteringTexture, functions.glsl:1285-1330, 1369-1383), each with its fused
// luminance / accumulation epilogue (atmosphere/model.cc:142-157, 192-208).
//
// Mapping: one block per (layer k, mu row j); one thread per x = i_nu * mu_s_n + i_mu_s. All
// texels of a block share the ray (r, mu): the 51 sample points, their radii, the transmittance
// along the ray and the (r, mu) interpolation footprint in the source table are identical for
// every thread. So per sample:
//   phase A (once per block, fp64): 51 threads compute the sample geometry and the table taps of
//           the shared axes; then all threads compute the path transmittance, one (sample, channel)
//           pair each;
//   stage   (all threads): the shared-axis interpolation is applied ONCE to a whole table row:
//           the 4 (layer, row) corner rows of the channel-interleaved source table (16 KB each at
//           15 channels, contiguous) are read with 128-bit loads from L2, combined in registers
//           and written to shared memory as one row of texels (multiple scattering); the 2
//           transmittance rows bracketing r_i likewise (single scattering). At the reference's row
//           width the corner rows live in register slots across samples: rows shared with the
//           previous sample are not read again, the next sample's rows are prefetched;
//   consume (per thread, fp32): only the thread-dependent axes remain (mu_s and nu, resp. the
//           sun-direction mu): 2-4 texel reads (128-bit) instead of 16 L2 gathers per channel.
// The stage buffer is double buffered: one __syncthreads per sample. Three kernels:
// multiple_scattering_rows_kernel (row width 256, more than 4 channels: the bench path),
// multiple_scattering_kernel (any width; also 256 with <= 4 channels, where a row is only 4 KB),
// single_scattering_kernel (both).
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

// Shared per-sample records (written in phase A, read by every warp once per sample as three
// 128-bit broadcast loads).
struct __align__(16) ScatterSample {   // multiple scattering
  // footprint of (r_i, mu_i) in the source table: row offsets (in texels) and bilinear weights
  int row00, row01, row10, row11;
  float w00, w01, w10, w11;
  float d;          // distance along the ray
  float inv_r;      // 1 / r_i
  float pad0, pad1;
};
struct __align__(16) SunSample {       // single scattering: sun-lookup geometry at r_i
  float d;          // distance along the ray
  float inv_r;      // 1 / r_i
  float q;          // (top - r_i)(top + r_i)
  float d_min;      // top - r_i
  float x_scale;    // (t_w - 1) / (d_max - d_min)
  float cos_h;      // cosine of the horizon angle at r_i (functions.glsl:556-557)
  float inv_sun_w;  // 1 / (2 sin_h alpha_s)
  float wy;         // weight of transmittance row y1
  float dens_r, dens_m;
  int y0, y1;       // transmittance rows bracketing r_i
};
template <typename S>
__device__ __forceinline__ S load_sample(const S* p) {
  static_assert(sizeof(S) == 48, "three 16-byte vectors");
  union { S s; float4 v[3]; } u;
  const float4* q = reinterpret_cast<const float4*>(p);
  u.v[0] = q[0]; u.v[1] = q[1]; u.v[2] = q[2];
  return u.s;
}

// Phase A, one thread per sample (fp64): fills the per-sample record and
// Tw[c] = T(r, mu, d_i)[c] * trapezoid weight * dx.
__device__ __forceinline__ void fill_sample(const PasGeometry& g, double d, double r_i, double mu_i,
                                            double rho_i, bool hit, ScatterSample* out) {
  ScatterSample s;
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
  s.d = (float)d;
  s.inv_r = (float)(1.0 / r_i);
  s.pad0 = s.pad1 = 0.f;
  *out = s;
}
__device__ __forceinline__ void fill_sample(const PasGeometry& g, double d, double r_i, double mu_i,
                                            double rho_i, bool hit, SunSample* out) {
  SunSample s;
  s.d = (float)d;
  s.inv_r = (float)(1.0 / r_i);
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
  *out = s;
}

// Transmittance taps of one sample: GetTransmittance(r, mu, d_i, hit) is the ratio of two bilinear
// fetches (functions.glsl:493-519); the taps are channel independent, the fetches are not.
struct PathTaps {
  Tap ax, ay, bx, by;
  double w;  // trapezoid weight * dx
};

// Phase A, step 1: one thread per sample (fp64 geometry): the per-sample record and the taps.
template <typename Sample>
__device__ void ray_sample_geometry(const PasGeometry& g, double r, double rho, double mu, bool hit,
                                    double d_end, int i, Sample* out, PathTaps* taps) {
  const double dx = d_end / PAS_RAY_SAMPLES;
  const double d = i * dx;
  const double r_i = d_clamp(sqrt(d * d + 2.0 * r * mu * d + r * r), g.bottom, g.top);
  const double mu_i = d_clamp((r * mu + d) / r_i, -1.0, 1.0);
  const double rho_i = sqrt(d_pos(r_i * r_i - g.bottom * g.bottom));
  fill_sample(g, d, r_i, mu_i, rho_i, hit, out);
  // GetTransmittance(r, mu, d, hit) (functions.glsl:493-519)
  double xa, ya, xb, yb;
  if (hit) {
    transmittance_xy(g, r_i, -mu_i, &xa, &ya);
    transmittance_xy(g, r, -mu, &xb, &yb);
  } else {
    transmittance_xy(g, r, mu, &xa, &ya);
    transmittance_xy(g, r_i, mu_i, &xb, &yb);
  }
  PathTaps t;
  t.ax = make_tap(xa, g.sz.t_w);
  t.ay = make_tap(ya, g.sz.t_h);
  t.bx = make_tap(xb, g.sz.t_w);
  t.by = make_tap(yb, g.sz.t_h);
  t.w = ((i == 0 || i == PAS_RAY_SAMPLES) ? 0.5 : 1.0) * dx;
  *taps = t;
}

// Phase A, step 2: all threads, one (sample, channel) pair each:
// Tw[i][c] = T(r, mu, d_i)[c] * trapezoid weight * dx. (A sample thread doing its NC channels one
// after the other kept the other 200 threads of the block waiting on ~15 dependent L2 round trips.)
template <int NC>
__device__ __forceinline__ void ray_sample_transmittance(const PasGeometry& g, const float* __restrict__ T,
                                                         const PathTaps* taps, float (*Tw)[PAS_CHANNEL_PITCH(NC)],
                                                         int tid, int nthreads) {
  constexpr int CP = PAS_CHANNEL_PITCH(NC);
  for (int idx = tid; idx < kSamples * CP; idx += nthreads) {
    const int i = idx / CP, c = idx % CP;
    float v = 0.f;
    if (c < NC) {
      const PathTaps& p = taps[i];
      const double t = fmin(fetch_t(T, CP, c, g.sz.t_w, p.ax, p.ay) / fetch_t(T, CP, c, g.sz.t_w, p.bx, p.by), 1.0);
      v = (float)(t * p.w);
    }
    Tw[i][c] = v;
  }
}

// Four 128-bit row loads (p, p + stride, ...) into a register slot, behind a block-uniform BRANCH on
// `bit`. Written as `if (bit) { loads }`, ptxas emits predicated loads, and a predicated-off LDG still
// takes an issue slot and one pass through the L1 data pipe (measured: the "misc" shared wavefronts of
// the ray-march kernels are their predicated-off row loads, 10 % of the pipe that bounds them).
// bra.uni keeps the branch; the destinations are read-write operands, so the loads land in the
// slot's own registers and nothing waits for them here.
template <int STRIDE_BYTES>
__device__ __forceinline__ void load_slot4(float4 (&R)[4], const float4* p, int bit) {
  asm volatile(
      "{\n"
      " .reg .pred take;\n"
      " setp.ne.b32 take, %16, 0;\n"
      " @!take bra.uni SKIP;\n"
      " bar.warp.sync 0xffffffff;\n"
      " ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%17];\n"
      " ld.global.nc.v4.f32 {%4, %5, %6, %7}, [%17 + %18];\n"
      " ld.global.nc.v4.f32 {%8, %9, %10, %11}, [%17 + %19];\n"
      " ld.global.nc.v4.f32 {%12, %13, %14, %15}, [%17 + %20];\n"
      "SKIP:\n"
      "}\n"
      : "+f"(R[0].x), "+f"(R[0].y), "+f"(R[0].z), "+f"(R[0].w), "+f"(R[1].x), "+f"(R[1].y), "+f"(R[1].z), "+f"(R[1].w),
        "+f"(R[2].x), "+f"(R[2].y), "+f"(R[2].z), "+f"(R[2].w), "+f"(R[3].x), "+f"(R[3].y), "+f"(R[3].z), "+f"(R[3].w)
      : "r"(bit), "l"(p), "n"(STRIDE_BYTES), "n"(2 * STRIDE_BYTES), "n"(3 * STRIDE_BYTES));
}

// RGBA store / accumulate into a final table (fp32 or fp16 texels). Returns the value of the texel
// after the accumulation, before the rounding of the store.
__device__ __forceinline__ float4 final_rgba(void* base, size_t texel, float4 v, int half, bool add) {
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
  return v;
}
// The same texel into a host-mapped copy of the table (FinalTables::host_scattering), same bits.
__device__ __forceinline__ void host_rgba(void* base, size_t texel, float4 v, int half) {
  if (half) {
    __half2 lo = __floats2half2_rn(v.x, v.y), hi = __floats2half2_rn(v.z, v.w);
    uint2 u;
    u.x = *reinterpret_cast<unsigned*>(&lo);
    u.y = *reinterpret_cast<unsigned*>(&hi);
    reinterpret_cast<uint2*>(base)[texel] = u;
  } else {
    reinterpret_cast<float4*>(base)[texel] = v;
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

// Staged rows live in shared memory as Q = CP / 4 planes of 16-byte vectors, plane q holding vector
// q of every texel: position (in float4 units) = q * pitch + x, pitch = W rounded up to 8, plus
// 8 / Q. While staging, the 8 lanes of a quarter warp write 8 / Q consecutive texels x Q planes: the
// plane padding sends them to 8 different 16-byte bank groups. While gathering, a lane reads the same
// plane at its own texel; neighbouring lanes read neighbouring texels. With the reference's row
// width (W = 256 = one texel per thread) pitch is a compile-time constant, so the plane offsets fold
// into the LDS/STS immediates and the staging address is loop invariant.
template <int Q>
__device__ __forceinline__ int plane_pitch(int w) { return ((w + 7) & ~7) + 8 / Q; }

// Rotation of x inside its aligned group of 8 by (x >> 3): texels read with a stride of 8 (sun
// lookups of neighbouring mu_s columns) then fall in different bank groups too.
__device__ __forceinline__ int rot8(int x) { return (x & ~7) | ((x + (x >> 3)) & 7); }

__device__ __forceinline__ float4 lerp4(float w, float4 a, float4 b) {
  return make_float4(fmaf(w, b.x - a.x, a.x), fmaf(w, b.y - a.y, a.y), fmaf(w, b.z - a.z, a.z),
                     fmaf(w, b.w - a.w, a.w));
}
__device__ __forceinline__ float4 combine4(float w0, float4 a, float w1, float4 b, float w2, float4 c,
                                           float w3, float4 d) {
  return make_float4(fmaf(w0, a.x, fmaf(w1, b.x, fmaf(w2, c.x, w3 * d.x))),
                     fmaf(w0, a.y, fmaf(w1, b.y, fmaf(w2, c.y, w3 * d.y))),
                     fmaf(w0, a.z, fmaf(w1, b.z, fmaf(w2, c.z, w3 * d.z))),
                     fmaf(w0, a.w, fmaf(w1, b.w, fmaf(w2, c.w, w3 * d.w))));
}
__device__ __forceinline__ void fma4(float4& acc, float4 v, float4 t) {
  acc.x = fmaf(v.x, t.x, acc.x);
  acc.y = fmaf(v.y, t.y, acc.y);
  acc.z = fmaf(v.z, t.z, acc.z);
  acc.w = fmaf(v.w, t.w, acc.w);
}

// ---- multiple scattering ----------------------------------------------------------------------
// WIDTH > 0: the row width nu_n * mu_s_n is known at compile time and equals the block size.
template <int NC, int MAXT, int MINB, int WIDTH>
__global__ void __launch_bounds__(MAXT, MINB)
multiple_scattering_kernel(const __grid_constant__ PasGeometry g,
                           const __grid_constant__ PasSpectrum sp, const float* __restrict__ T,
                           const float* __restrict__ dJ, float* __restrict__ dS, FinalTables fin,
                           int k_begin, int k_stride) {
  constexpr int CP = PAS_CHANNEL_PITCH(NC), Q = CP / 4;
  extern __shared__ __align__(16) float smem_dyn[];
  __shared__ ScatterSample sSample[kSamples];
  __shared__ __align__(16) float sTw[kSamples][CP];

  const int tid = threadIdx.x;
  const int j = blockIdx.x, k = k_begin + blockIdx.y * k_stride;
  const int mu_n = g.sz.mu_n, nu_n = g.sz.nu_n, mu_s_n = g.sz.mu_s_n;
  const int width = WIDTH > 0 ? WIDTH : nu_n * mu_s_n;
  const int nthreads = WIDTH > 0 ? WIDTH : (int)blockDim.x;
  const int pitch = plane_pitch<Q>(width);
  float4* sRow = reinterpret_cast<float4*>(smem_dyn);  // [2][Q * pitch]
  const float4* dJ4 = reinterpret_cast<const float4*>(dJ);

  __shared__ PathTaps sTaps[kSamples];
  const BlockRay ray = block_ray(g, k, j);
  if (tid < kSamples) {
    ray_sample_geometry(g, ray.r, ray.rho, ray.mu, ray.hit, ray.d_end, tid, &sSample[tid], &sTaps[tid]);
  }
  __syncthreads();
  ray_sample_transmittance<NC>(g, T, sTaps, sTw, tid, (int)blockDim.x);

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
  // Texels whose nu lies exactly on a slab (every texel whose nu is not clamped, about 70 %) need 2
  // instead of 4 texel reads per sample; consecutive texels share their slab and mostly their clamping,
  // so whole warps qualify without sorting the row (the rows kernel below sorts it).
  const bool on_slab = !active || wnu == 0.0f || wnu == 1.0f || tnu.i0 == tnu.i1;
  const bool warp_on_slab = __all_sync(0xffffffffu, on_slab);
  const int slab_s = wnu == 1.0f ? slab1 : slab0;
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
      const ScatterSample s = load_sample(&sSample[i]);
      float4* buf = sRow + (i & 1) * Q * pitch;
      // stage: bilinear in (r, mu) applied to the whole row of texels. The row is walked as a flat
      // array of 16-byte vectors so that a warp reads 512 contiguous bytes per load; all the loads
      // of a thread are issued before the first use.
      {
        const float4* p00 = dJ4 + (size_t)s.row00 * Q;
        const float4* p01 = dJ4 + (size_t)s.row01 * Q;
        const float4* p10 = dJ4 + (size_t)s.row10 * Q;
        const float4* p11 = dJ4 + (size_t)s.row11 * Q;
        const int nvec = width * Q;
        if (WIDTH > 0) {
          // flat vector f = tid + it * WIDTH: plane f % Q = tid % Q, texel f / Q = tid / Q + it * (WIDTH / Q)
          float4* dst = buf + (tid % Q) * pitch + tid / Q;
          float4 a[Q], b[Q], c[Q], d[Q];
#pragma unroll
          for (int it = 0; it < Q; ++it) {
            const int f = tid + it * WIDTH;
            a[it] = __ldg(p00 + f); b[it] = __ldg(p01 + f); c[it] = __ldg(p10 + f); d[it] = __ldg(p11 + f);
          }
#pragma unroll
          for (int it = 0; it < Q; ++it) {
            dst[it * (WIDTH / Q)] = combine4(s.w00, a[it], s.w01, b[it], s.w10, c[it], s.w11, d[it]);
          }
        } else {
          for (int f = tid; f < nvec; f += nthreads) {
            buf[(f % Q) * pitch + f / Q] =
                combine4(s.w00, __ldg(p00 + f), s.w01, __ldg(p01 + f), s.w10, __ldg(p10 + f), s.w11, __ldg(p11 + f));
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
        const float4* tw4 = reinterpret_cast<const float4*>(sTw[i]);
        if (warp_on_slab) {
          const float4* a0 = buf + slab_s + tm.i0;
          const float4* a1 = buf + slab_s + tm.i1;
          const float w0 = 1.0f - wm;
#pragma unroll
          for (int q = 0; q < Q; ++q) {
            const float4 u = a0[q * pitch], v = a1[q * pitch];
            fma4(acc[q], make_float4(fmaf(w0, u.x, wm * v.x), fmaf(w0, u.y, wm * v.y), fmaf(w0, u.z, wm * v.z),
                                     fmaf(w0, u.w, wm * v.w)), tw4[q]);
          }
        } else {
          // the four corner weights of the (mu_s, nu) bilinear fetch
          const float w11 = wnu * wm, w10 = wnu - w11, w01 = wm - w11, w00 = 1.0f - wnu - wm + w11;
          const float4* a0 = buf + slab0 + tm.i0;
          const float4* a1 = buf + slab0 + tm.i1;
          const float4* b0 = buf + slab1 + tm.i0;
          const float4* b1 = buf + slab1 + tm.i1;
#pragma unroll
          for (int q = 0; q < Q; ++q) {
            fma4(acc[q], combine4(w00, a0[q * pitch], w01, a1[q * pitch], w10, b0[q * pitch], w11, b1[q * pitch]),
                 tw4[q]);
          }
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
  const float4 total = final_rgba(fin.scattering, texel,
                                  make_float4(rgb[0] * inv_pr, rgb[1] * inv_pr, rgb[2] * inv_pr, 0.f),
                                  fin.half_precision, true);
  // consecutive threads own consecutive texels: the host copy is written in full 32-byte sectors
  if (fin.host_scattering != nullptr) host_rgba(fin.host_scattering, texel, total, fin.half_precision);
}

// ---- per-ray setup tables ------------------------------------------------------------------------------
// Everything a ray-march block computes before its sample loop depends on the ray (layer k, mu row j)
// and on the transmittance table only -- not on the scattering order: the 51 sample records, the path
// transmittances Tw[sample][channel], the on-slab permutation of the row's texels. Computed inside the
// kernels it is a latency-bound prologue (fp64 geometry on 51 threads, then 51 x 16 ratios of bilinear
// fp64 fetches) that costs 0.19 of the 0.94 ms of a multiple-scattering pass. ray_setup_kernel runs
// the SAME code once per Init and channel group and leaves the results in HBM (9 KB per ray, 36 MB);
// the single-scattering pass and every multiple-scattering pass then start by copying them in.
struct RaySetupLayout {
  size_t plan, sun, tw, perm, stride;   // byte offsets inside a ray's record, record size
};
__host__ __device__ inline RaySetupLayout ray_setup_layout(int cp) {
  RaySetupLayout l;
  l.plan = 0;
  l.sun = l.plan + kSamples * 48;
  l.tw = l.sun + kSamples * 48;
  l.perm = l.tw + (size_t)kSamples * cp * sizeof(float);
  l.stride = (l.perm + 256 * sizeof(int) + 15) & ~(size_t)15;
  return l;
}
// block-wide copy of n 16-byte vectors
__device__ __forceinline__ void copy16(void* dst, const void* src, int n, int tid, int nthreads) {
  for (int i = tid; i < n; i += nthreads) reinterpret_cast<float4*>(dst)[i] = reinterpret_cast<const float4*>(src)[i];
}

// ---- multiple scattering, reference row width (256 texels = block size) -------------------------
// Same mapping as above, with three changes that cut the L1 data-pipe wavefronts (the bound of this
// kernel):
//  * row slots: consecutive samples of a ray share on average 2.5 of their 4 (layer, mu) corner rows.
//    Every thread keeps its 16-byte vectors of four rows in registers ("slots"); a per-block plan
//    (built once, sequentially, by thread 0) says for every sample which row each slot must hold,
//    with which weight, and which slots have to be (re)loaded. Rows that stay are not read again;
//    the loads of sample i + 1 are issued right after the combine of sample i and land while the
//    block gathers sample i.
//  * texels whose nu lies exactly on a slab (about 70 %: every texel whose nu is not clamped,
//    functions.glsl:923-925) need 2 instead of 4 texel reads per sample. The block's texels are
//    permuted so that those come first and whole warps take the short path (bit-identical: the
//    skipped terms are exact zeros).
//  * staged rows split by texel parity: a lane reads the two neighbouring mu_s texels (i0, i0 + 1) of a
//    slab with two 128-bit loads, and a quarter warp (the unit the shared-memory pipe serves per
//    wavefront) holds 8 consecutive columns whose i0 spread over 8 s texels, s = the local stretch
//    of the mu_s map between the texel and the sample (up to ~2). With the texels of a plane in
//    their natural order two lanes hit the same 16-byte bank group as soon as s > 1 (1.23 extra
//    wavefronts per load, measured, and reproduced by tools/probe/ms_bank_model.py). Here a plane
//    holds the even texels of slab n at positions [n P, n P + mu_s_n / 2) and the odd ones HALF
//    further; one load takes the even texel of every lane's pair, the other the odd one: 8 lanes
//    then cover 8 s / 2 consecutive positions and stay conflict-free up to s = 2. The slab pitch
//    P = mu_s_n / 2 + 1 (odd) separates the two ends of a quarter warp that straddles two slabs
//    (model: 5.23 wavefronts per 128-bit load in the natural order, 4.31 with the parity split and
//    the lane predicate below, 3.96 with the odd slab pitch; 4 = no conflict, no broadcast).
//    HALF = 4 (mod 8) and pitch = 4 / Q (mod 8) keep the staging stores of a quarter warp (8 / Q
//    consecutive texels x Q planes) on 8 different bank groups.
//    Row shape taken: the reference's 8 x 32 (rows_shape_ok), so that every offset is a constant.
constexpr int kRowsHalf = 128 + 8 + 4;   // 256 / 2 texels + one pad per slab, = 4 (mod 8)
template <int Q>
__host__ __device__ constexpr int rows_pitch() { return 2 * kRowsHalf + (Q > 1 ? 4 / Q : 0); }
inline bool rows_shape_ok(const PasGeometry& g) {
  return g.sz.nu_n == 8 && g.sz.mu_s_n == 32;
}

struct __align__(16) SlotSample {
  int row[4];       // row offset (in texels) held by each slot during this sample
  float w[4];       // weight of each slot
  float d, inv_r;
  int mask;         // slots to load before this sample's combine
  int pad;
};

// Prologue of a multiple-scattering block for the ray (layer k, mu row j): leaves the slot plan in
// sSample (as SlotSample), the path transmittances in sTw and the on-slab permutation in sPerm.
// (WITH_TW = false: the caller already holds the path transmittances, which do not depend on the pass.)
template <int NC, bool WITH_TW = true>
__device__ void rows_prologue(const PasGeometry& g, const float* __restrict__ T, const BlockRay& ray, int tid,
                              ScatterSample* sSample, float (*sTw)[PAS_CHANNEL_PITCH(NC)], int* sPerm,
                              int* sCount) {
  constexpr int WIDTH = 256;
  const int mu_n = g.sz.mu_n, nu_n = g.sz.nu_n, mu_s_n = g.sz.mu_s_n;
  __shared__ PathTaps sTaps[kSamples];
  if (tid < kSamples) {
    ray_sample_geometry(g, ray.r, ray.rho, ray.mu, ray.hit, ray.d_end, tid, &sSample[tid], &sTaps[tid]);
  }

  // ---- permutation: texels with nu on a slab first ---------------------------------------------
  {
    const int i_nu = tid / mu_s_n, i_mu_s = tid % mu_s_n;
    const double nu_d = scattering_slab_nu(g, i_nu, ray.mu, scattering_col_mu_s(g, i_mu_s));
    const Tap t = make_tap((nu_d + 1.0) * 0.5 * (nu_n - 1), nu_n);
    const bool on_slab = t.w == 0.0f || t.w == 1.0f || t.i0 == t.i1;
    const unsigned ballot = __ballot_sync(0xffffffffu, on_slab);
    const int lane = tid & 31, warp = tid >> 5;
    if (lane == 0) sCount[warp] = __popc(ballot);
    __syncthreads();
    int before = 0, total = 0;
#pragma unroll
    for (int w = 0; w < WIDTH / 32; ++w) {
      const int c = sCount[w];
      before += w < warp ? c : 0;
      total += c;
    }
    const int rank_in_warp = __popc(ballot & ((1u << lane) - 1u));
    // on-slab texels keep their relative order in [0, total); the others follow in [total, WIDTH)
    const int pos = on_slab ? before + rank_in_warp : total + (warp * 32 - before) + (lane - rank_in_warp);
    sPerm[pos] = tid;
  }
  // (the barrier inside the permutation published the taps of phase A)
  if (WITH_TW) ray_sample_transmittance<NC>(g, T, sTaps, sTw, tid, WIDTH);
  // ---- slot plan ---------------------------------------------------------------------------------
  // Row (layer k, mu row j) always lives in slot 2 (k & 1) + (j & 1): the four corner rows of a sample
  // fall in four different slots, and a row shared with the previous sample is found where it was.
  __syncthreads();
  SlotSample mine;
  if (tid < kSamples) {
    const ScatterSample s = load_sample(&sSample[tid]);
    const int rows[4] = {s.row00, s.row01, s.row10, s.row11};
    const float ws[4] = {s.w00, s.w01, s.w10, s.w11};
#pragma unroll
    for (int t = 0; t < 4; ++t) { mine.row[t] = -1; mine.w[t] = 0.f; }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int rj = rows[c] / WIDTH;              // k * mu_n + j (mu_n is even)
      const int slot = 2 * ((rj / mu_n) & 1) + (rj & 1);
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        if (slot == t) { mine.row[t] = rows[c]; mine.w[t] += ws[c]; }  // clamped taps name a row twice
      }
    }
    mine.d = s.d;
    mine.inv_r = s.inv_r;
    mine.pad = 0;
  }
  __syncthreads();  // every sample has been read in its first form
  SlotSample* plan_w = reinterpret_cast<SlotSample*>(sSample);
  if (tid < kSamples) {
    mine.mask = 0;
    plan_w[tid] = mine;
  }
  __syncthreads();
  if (tid < kSamples) {
    // a slot is (re)loaded when the previous sample left another row in it (or did not use it); a
    // slot this sample does not use (row -1, weight 0) keeps its registers, whatever they hold
    int mask = 0;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int prev = tid > 0 ? plan_w[tid - 1].row[t] : -1;
      if (mine.row[t] >= 0 && mine.row[t] != prev) mask |= 1 << t;
    }
    plan_w[tid].mask = mask;
  }
  __syncthreads();
}

template <int NC>
__global__ void __launch_bounds__(256, 2)
multiple_scattering_rows_kernel(const __grid_constant__ PasGeometry g,
                                const __grid_constant__ PasSpectrum sp, const float* __restrict__ T,
                                const float* __restrict__ dJ, float* __restrict__ dS,
                                FinalTables fin, int k_begin, int k_stride, const char* __restrict__ setup) {
  constexpr int CP = PAS_CHANNEL_PITCH(NC), Q = CP / 4, WIDTH = 256;
  static_assert(sizeof(SlotSample) == sizeof(ScatterSample), "plan is rewritten in place");
  extern __shared__ __align__(16) float smem_dyn[];
  __shared__ ScatterSample sSample[kSamples];
  __shared__ __align__(16) float sTw[kSamples][CP];
  __shared__ int sPerm[WIDTH];
  __shared__ int sCount[WIDTH / 32];

  const int tid = threadIdx.x;
  const int j = blockIdx.x, k = k_begin + blockIdx.y * k_stride;
  const int mu_n = g.sz.mu_n;
  constexpr int nu_n = 8, mu_s_n = 32;   // rows_shape_ok
  constexpr int HALF = kRowsHalf, pitch = rows_pitch<Q>();
  constexpr int slab_pitch = mu_s_n / 2 + 1;
  float4* sRow = reinterpret_cast<float4*>(smem_dyn);  // [2][Q * pitch]
  const float4* dJ4 = reinterpret_cast<const float4*>(dJ);

  const BlockRay ray = block_ray(g, k, j);
  if (setup != nullptr) {
    // the prologue was run once for all passes by ray_setup_kernel
    const RaySetupLayout lay = ray_setup_layout(CP);
    const char* rec = setup + ((size_t)k * mu_n + j) * lay.stride;
    copy16(sSample, rec + lay.plan, kSamples * 3, tid, WIDTH);
    copy16(sTw, rec + lay.tw, kSamples * CP / 4, tid, WIDTH);
    copy16(sPerm, rec + lay.perm, WIDTH / 4, tid, WIDTH);
    __syncthreads();
  } else {
    rows_prologue<NC>(g, T, ray, tid, sSample, sTw, sPerm, sCount);
  }
  const SlotSample* plan = reinterpret_cast<const SlotSample*>(sSample);

  // per-thread axes: mu_s (column) and nu (slab) of the texel this thread owns
  const int x = sPerm[tid];
  const int i_nu = x / mu_s_n, i_mu_s = x % mu_s_n;
  const double mu_s_d = scattering_col_mu_s(g, i_mu_s);
  const double nu_d = scattering_slab_nu(g, i_nu, ray.mu, mu_s_d);
  const float nu = (float)nu_d;
  const float r_mu_s = (float)(ray.r * mu_s_d);
  const float bottom = (float)g.bottom;
  const Tap tnu = make_tap((nu_d + 1.0) * 0.5 * (nu_n - 1), nu_n);
  const float wnu = tnu.w;
  const bool on_slab = wnu == 0.0f || wnu == 1.0f || tnu.i0 == tnu.i1;
  const bool warp_on_slab = __all_sync(0xffffffffu, on_slab);
  // plane positions of the slabs; a lane on a slab takes that slab alone with weight 1, the others
  // take slab0 and slab1
  const int pos_a = (on_slab && wnu == 1.0f ? tnu.i1 : tnu.i0) * slab_pitch, pos_b = tnu.i1 * slab_pitch;
  const float w_a = on_slab ? 1.0f : 1.0f - wnu, w_b = wnu;
  MuSMap map;
  map.H2 = (float)(g.H * g.H);
  map.d_min = (float)(g.top - g.bottom);
  map.inv_range = (float)(1.0 / (g.H - (g.top - g.bottom)));
  map.inv_A = (float)(1.0 / g.mus_A);
  map.scale = (float)(mu_s_n - 1);

  float4 acc[Q];
#pragma unroll
  for (int q = 0; q < Q; ++q) acc[q] = make_float4(0.f, 0.f, 0.f, 0.f);

  if (ray.d_end > 0.0) {
    float4 R0[Q], R1[Q], R2[Q], R3[Q];
#pragma unroll
    for (int it = 0; it < Q; ++it) R0[it] = R1[it] = R2[it] = R3[it] = make_float4(0.f, 0.f, 0.f, 0.f);
    // flat vector f = tid + it * WIDTH of a row: plane f % Q = tid % Q, texel f / Q = tid / Q + it * (WIDTH / Q)
    // (WIDTH / Q is a multiple of mu_s_n: the texels of a thread share their column, WIDTH / Q / mu_s_n
    // slabs apart)
    const int col0 = (tid / Q) % mu_s_n;
    float4* const dst0 = sRow + (tid % Q) * pitch + (tid / Q) / mu_s_n * slab_pitch + (col0 >> 1) + (col0 & 1) * HALF;
    const int dst_step = WIDTH / Q / mu_s_n * slab_pitch;
#define PAS_LOAD_SLOT(R, S, T)                                                           \
      if (Q == 4) {                                                                      \
        load_slot4<WIDTH * 16>(reinterpret_cast<float4(&)[4]>(R), dJ4 + (size_t)(S).row[T] * Q + tid, (S).mask & (1 << T)); \
      } else if ((S).mask & (1 << T)) {                                                  \
        const float4* p = dJ4 + (size_t)(S).row[T] * Q + tid;                            \
        _Pragma("unroll") for (int it = 0; it < Q; ++it) R[it] = __ldg(p + it * WIDTH);  \
      }
#define PAS_LOAD_SLOTS(S)                                                                \
    {                                                                                    \
      PAS_LOAD_SLOT(R0, S, 0)                                                            \
      PAS_LOAD_SLOT(R1, S, 1)                                                            \
      PAS_LOAD_SLOT(R2, S, 2)                                                            \
      PAS_LOAD_SLOT(R3, S, 3)                                                            \
    }
    SlotSample s = load_sample(&plan[0]);
    PAS_LOAD_SLOTS(s)
    for (int i = 0; i < kSamples; ++i) {
      float4* buf = sRow + (i & 1) * Q * pitch;
      float4* dst = dst0 + (i & 1) * Q * pitch;
      {
        // the path transmittance (x trapezoid weight x dx) of the sample multiplies every texel of the
        // row alike: applied here, once per staged vector, instead of once per gathered texel
        const float4 tw = reinterpret_cast<const float4*>(sTw[i])[tid % Q];
#pragma unroll
        for (int it = 0; it < Q; ++it) {
          const float4 v = combine4(s.w[0], R0[it], s.w[1], R1[it], s.w[2], R2[it], s.w[3], R3[it]);
          dst[it * dst_step] = make_float4(v.x * tw.x, v.y * tw.y, v.z * tw.z, v.w * tw.w);
        }
      }
      const float s_d = s.d, s_inv_r = s.inv_r;
      if (i + 1 < kSamples) {
        s = load_sample(&plan[i + 1]);
        PAS_LOAD_SLOTS(s)
      }
      __syncthreads();
      // mu_s at the sample: (r mu_s + d nu) / r_i (functions.glsl:1314)
      const float mu_s_i = f_clamp(fmaf(s_d, nu, r_mu_s) * s_inv_r, -1.0f, 1.0f);
      const float xs = f_clamp(f_mu_s_texel_x(map, bottom * mu_s_i), 0.0f, map.scale);
      const Tap tm = make_tap_f(xs, mu_s_n);   // i0 <= mu_s_n - 2, i1 = i0 + 1
      // the even and the odd texel of the pair (i0, i0 + 1), with their weights
      const bool odd0 = (tm.i0 & 1) != 0;
      const float4* pe = buf + ((tm.i0 + 1) >> 1);
      const float4* po = buf + HALF + (tm.i0 >> 1);
      const float we = odd0 ? tm.w : 1.0f - tm.w, wo = odd0 ? 1.0f - tm.w : tm.w;
      if (warp_on_slab) {
        const float4* ae = pe + pos_a;
        const float4* ao = po + pos_a;
#pragma unroll
        for (int q = 0; q < Q; ++q) {
          const float4 u = ae[q * pitch], v = ao[q * pitch];
          acc[q].x = fmaf(we, u.x, fmaf(wo, v.x, acc[q].x));
          acc[q].y = fmaf(we, u.y, fmaf(wo, v.y, acc[q].y));
          acc[q].z = fmaf(we, u.z, fmaf(wo, v.z, acc[q].z));
          acc[q].w = fmaf(we, u.w, fmaf(wo, v.w, acc[q].w));
        }
      } else {
        // lanes on a slab read it alone: a quarter warp of them costs no wavefront on the second slab
        {
          const float4* ae = pe + pos_a;
          const float4* ao = po + pos_a;
          const float wae = w_a * we, wao = w_a * wo;
#pragma unroll
          for (int q = 0; q < Q; ++q) {
            const float4 u = ae[q * pitch], v = ao[q * pitch];
            acc[q].x = fmaf(wae, u.x, fmaf(wao, v.x, acc[q].x));
            acc[q].y = fmaf(wae, u.y, fmaf(wao, v.y, acc[q].y));
            acc[q].z = fmaf(wae, u.z, fmaf(wao, v.z, acc[q].z));
            acc[q].w = fmaf(wae, u.w, fmaf(wao, v.w, acc[q].w));
          }
        }
        if (!on_slab) {
          const float4* be = pe + pos_b;
          const float4* bo = po + pos_b;
          const float wbe = w_b * we, wbo = w_b * wo;
#pragma unroll
          for (int q = 0; q < Q; ++q) {
            const float4 u = be[q * pitch], v = bo[q * pitch];
            acc[q].x = fmaf(wbe, u.x, fmaf(wbo, v.x, acc[q].x));
            acc[q].y = fmaf(wbe, u.y, fmaf(wbo, v.y, acc[q].y));
            acc[q].z = fmaf(wbe, u.z, fmaf(wbo, v.z, acc[q].z));
            acc[q].w = fmaf(wbe, u.w, fmaf(wbo, v.w, acc[q].w));
          }
        }
      }
    }
#undef PAS_LOAD_SLOTS
#undef PAS_LOAD_SLOT
  }
  const size_t texel = ((size_t)k * mu_n + j) * WIDTH + x;
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
  const float4 total = final_rgba(fin.scattering, texel,
                                  make_float4(rgb[0] * inv_pr, rgb[1] * inv_pr, rgb[2] * inv_pr, 0.f),
                                  fin.half_precision, true);
  if (fin.host_scattering != nullptr) {
    // Host copy of the row: the threads own the texels in permuted order, so the row goes through
    // shared memory and leaves as 16-byte vectors in texel order (full PCIe write payloads).
    __syncthreads();  // the staged rows are dead
    const size_t row = ((size_t)k * mu_n + j) * WIDTH;
    if (fin.half_precision) {
      host_rgba(smem_dyn, (size_t)x, total, 1);
      __syncthreads();
      if (tid < WIDTH / 2) {
        reinterpret_cast<uint4*>(reinterpret_cast<uint2*>(fin.host_scattering) + row)[tid] =
            reinterpret_cast<const uint4*>(smem_dyn)[tid];
      }
    } else {
      host_rgba(smem_dyn, (size_t)x, total, 0);
      __syncthreads();
      (reinterpret_cast<float4*>(fin.host_scattering) + row)[tid] = reinterpret_cast<const float4*>(smem_dyn)[tid];
    }
  }
}

// ---- single scattering ------------------------------------------------------------------------
// WIDTH > 0: nu_n * mu_s_n == t_w == WIDTH == block size.
// NU_LANES (WIDTH > 0 only): consecutive threads own the nu slabs of one mu_s column instead of the
// mu_s columns of one slab. The sun lookups of neighbouring lanes then fall on the same or on
// neighbouring transmittance texels (they differ by d nu / r_i only), which the 128-bit shared loads
// serve without bank conflicts; the staged row is stored unrotated.
// Prologue of a single-scattering block for the ray (layer k, mu row j): the sample records in sSample
// (with the transmittance-row slots to load in bits 16.. of y0 when the rows live in register slots) and
// the path transmittances in sTw. Ends without a barrier: the caller's next one publishes the results.
template <int NC, int WIDTH>
__device__ void sun_prologue(const PasGeometry& g, const float* __restrict__ T, const BlockRay& ray, int tid,
                             SunSample* sSample, float (*sTw)[PAS_CHANNEL_PITCH(NC)]) {
  __shared__ PathTaps sTaps[kSamples];
  if (tid < kSamples) {
    ray_sample_geometry(g, ray.r, ray.rho, ray.mu, ray.hit, ray.d_end, tid, &sSample[tid], &sTaps[tid]);
  }
  __syncthreads();
  ray_sample_transmittance<NC>(g, T, sTaps, sTw, tid, (int)blockDim.x);
  // Reference row width (WIDTH > 0): the two transmittance rows bracketing r_i live in two register
  // slots, even rows in A and odd rows in B. A row shared with the previous sample is not read again
  // (about half of them), and the rows of sample i + 1 are requested right after the staged row of
  // sample i is written, so that they travel while the block does the sun lookups of sample i.
  // Bits 16.. of y0 carry the slots to (re)load: 1 = A, 2 = B.
  if (WIDTH > 0) {
    int mask = 0;
    if (tid < kSamples) {
      const int y0 = sSample[tid].y0, y1 = sSample[tid].y1;
      const int even = (y0 & 1) == 0 ? y0 : ((y1 & 1) == 0 ? y1 : -1);
      const int odd = (y0 & 1) == 1 ? y0 : ((y1 & 1) == 1 ? y1 : -1);
      int p_even = -1, p_odd = -1;
      if (tid > 0) {
        const int q0 = sSample[tid - 1].y0, q1 = sSample[tid - 1].y1;
        p_even = (q0 & 1) == 0 ? q0 : ((q1 & 1) == 0 ? q1 : -1);
        p_odd = (q0 & 1) == 1 ? q0 : ((q1 & 1) == 1 ? q1 : -1);
      }
      mask = (even >= 0 && even != p_even ? 1 : 0) | (odd >= 0 && odd != p_odd ? 2 : 0);
    }
    __syncthreads();
    if (tid < kSamples) sSample[tid].y0 |= mask << 16;
  }

}

template <int NC, int MAXT, int MINB, int WIDTH, bool NU_LANES>
__global__ void __launch_bounds__(MAXT, MINB)
single_scattering_kernel(const __grid_constant__ PasGeometry g,
                         const __grid_constant__ PasSpectrum sp, const float* __restrict__ T,
                         float* __restrict__ dR, float* __restrict__ dM, FinalTables fin,
                         int k_begin, int k_stride, const char* __restrict__ setup) {
  constexpr int CP = PAS_CHANNEL_PITCH(NC), Q = CP / 4;
  extern __shared__ __align__(16) float smem_dyn[];
  __shared__ SunSample sSample[kSamples];
  __shared__ __align__(16) float sTw[kSamples][CP];

  const int tid = threadIdx.x;
  const int j = blockIdx.x, k = k_begin + blockIdx.y * k_stride;
  const int mu_n = g.sz.mu_n, nu_n = g.sz.nu_n, mu_s_n = g.sz.mu_s_n;
  const int t_w = WIDTH > 0 ? WIDTH : g.sz.t_w;
  const int width = WIDTH > 0 ? WIDTH : nu_n * mu_s_n;
  const int nthreads = WIDTH > 0 ? WIDTH : (int)blockDim.x;
  // NU_LANES: the staged row is split by texel parity like the rows of multiple scattering (see
  // kRowsHalf): even texels at [0, WIDTH / 2), odd texels HALF further, one load per parity
  constexpr int HALF = WIDTH / 2 + 4;
  const int pitch = NU_LANES ? 2 * HALF + (Q > 1 ? 4 / Q : 0) : plane_pitch<Q>(t_w);
  float4* sRow = reinterpret_cast<float4*>(smem_dyn);  // [2][Q * pitch]
  const float4* T4 = reinterpret_cast<const float4*>(T);

  const BlockRay ray = block_ray(g, k, j);
  if (WIDTH > 0 && setup != nullptr) {
    // the prologue was run once for all passes by ray_setup_kernel
    const RaySetupLayout lay = ray_setup_layout(CP);
    const char* rec = setup + ((size_t)k * mu_n + j) * lay.stride;
    copy16(sSample, rec + lay.sun, kSamples * 3, tid, nthreads);
    copy16(sTw, rec + lay.tw, kSamples * CP / 4, tid, nthreads);
  } else {
    sun_prologue<NC, WIDTH>(g, T, ray, tid, sSample, sTw);
  }
  const int x = NU_LANES ? (tid % nu_n) * mu_s_n + tid / nu_n : tid;
  const bool active = x < width;
  const int i_nu = active ? x / mu_s_n : 0, i_mu_s = active ? x % mu_s_n : 0;
  const double mu_s_d = scattering_col_mu_s(g, i_mu_s);
  const double nu_d = scattering_slab_nu(g, i_nu, ray.mu, mu_s_d);
  const float nu = (float)nu_d;
  const float r_mu_s = (float)(ray.r * mu_s_d);
  const float x_max = (float)(t_w - 1);
  auto rot = [](int v) { return NU_LANES ? (v >> 1) + (v & 1) * HALF : rot8(v); };

  float4 accR[Q], accM[Q];
#pragma unroll
  for (int q = 0; q < Q; ++q) accR[q] = accM[q] = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncthreads();

  if (ray.d_end > 0.0) {
    float4 A[Q], B[Q];
#pragma unroll
    for (int it = 0; it < Q; ++it) A[it] = B[it] = make_float4(0.f, 0.f, 0.f, 0.f);
    SunSample s = load_sample(&sSample[0]);
#define PAS_LOAD_T_SLOTS(S)                                                                          \
    {                                                                                                \
      const int y0_ = (S).y0 & 0xffff, y1_ = (S).y1, m_ = (S).y0 >> 16;                              \
      const float4* pa_ = T4 + (size_t)((y0_ & 1) == 0 ? y0_ : y1_) * WIDTH * Q + tid;               \
      const float4* pb_ = T4 + (size_t)((y0_ & 1) == 1 ? y0_ : y1_) * WIDTH * Q + tid;               \
      if (Q == 4) {                                                                                  \
        load_slot4<WIDTH * 16>(reinterpret_cast<float4(&)[4]>(A), pa_, m_ & 1);                      \
        load_slot4<WIDTH * 16>(reinterpret_cast<float4(&)[4]>(B), pb_, m_ & 2);                      \
      } else {                                                                                       \
        if (m_ & 1) { _Pragma("unroll") for (int it = 0; it < Q; ++it) A[it] = __ldg(pa_ + it * WIDTH); } \
        if (m_ & 2) { _Pragma("unroll") for (int it = 0; it < Q; ++it) B[it] = __ldg(pb_ + it * WIDTH); } \
      }                                                                                              \
    }
    if (WIDTH > 0) PAS_LOAD_T_SLOTS(s)
    for (int i = 0; i < kSamples; ++i) {
      if (WIDTH == 0) s = load_sample(&sSample[i]);
      float4* buf = sRow + (i & 1) * Q * pitch;
      // stage the transmittance row at r_i (lerp of the two bracketing rows), flat 16-byte vectors
      if (WIDTH > 0) {
        // weights of the even / odd slot: y0 carries 1 - wy, y1 carries wy (wy = 0 when y0 == y1)
        const bool y0_odd = (s.y0 & 1) != 0;
        const float w_a = y0_odd ? s.wy : 1.0f - s.wy, w_b = y0_odd ? 1.0f - s.wy : s.wy;
#pragma unroll
        for (int it = 0; it < Q; ++it) {
          const float4 a = A[it], b = B[it];
          buf[(tid % Q) * pitch + rot(tid / Q + it * (WIDTH / Q))] =
              make_float4(fmaf(w_a, a.x, w_b * b.x), fmaf(w_a, a.y, w_b * b.y), fmaf(w_a, a.z, w_b * b.z),
                          fmaf(w_a, a.w, w_b * b.w));
        }
      } else {
        const float4* pa = T4 + (size_t)s.y0 * t_w * Q;
        const float4* pb = T4 + (size_t)s.y1 * t_w * Q;
        for (int f = tid; f < t_w * Q; f += nthreads) {
          buf[(f % Q) * pitch + rot(f / Q)] = lerp4(s.wy, __ldg(pa + f), __ldg(pb + f));
        }
      }
      const SunSample cur = s;
      if (WIDTH > 0 && i + 1 < kSamples) {
        s = load_sample(&sSample[i + 1]);
        PAS_LOAD_T_SLOTS(s)
      }
      __syncthreads();
      if (active) {
        // sun direction at the sample: r_i mu_s_i = r mu_s + d nu (functions.glsl:657)
        const float r_i = f_rcp(cur.inv_r);
        const float p = f_clamp(fmaf(cur.d, nu, r_mu_s), -r_i, r_i);
        // GetTransmittanceToSun (functions.glsl:552-563): table x from the distance to the top
        const float xt = f_clamp((f_dist_top(p, cur.q) - cur.d_min) * cur.x_scale, 0.0f, x_max);
        const Tap tu = make_tap_f(xt, t_w);
        const float mu_s_i = p * cur.inv_r;
        const float sm = f_sat(fmaf(mu_s_i - cur.cos_h, cur.inv_sun_w, 0.5f));
        const float vis = sm * sm * fmaf(-2.0f, sm, 3.0f);
        const float wr = vis * cur.dens_r, wm = vis * cur.dens_m;
        // NU_LANES: t0 = the even texel of the pair (i0, i0 + 1), t1 = the odd one
        const bool swap = NU_LANES && (tu.i0 & 1) != 0;
        const float4* t0 = buf + rot(swap ? tu.i1 : tu.i0);
        const float4* t1 = buf + rot(swap ? tu.i0 : tu.i1);
        const float w_t = swap ? 1.0f - tu.w : tu.w;
        // (folding the path transmittance into the staged row, as multiple scattering does, was measured
        // 1 % slower here: the staging sits on the critical path in front of the barrier)
        const float4* tw4 = reinterpret_cast<const float4*>(sTw[i]);
#pragma unroll
        for (int q = 0; q < Q; ++q) {
          const float4 tv = lerp4(w_t, t0[q * pitch], t1[q * pitch]);
          const float4 t = tw4[q];
          const float4 v = make_float4(tv.x * t.x, tv.y * t.y, tv.z * t.z, tv.w * t.w);
          fma4(accR[q], v, make_float4(wr, wr, wr, wr));
          fma4(accM[q], v, make_float4(wm, wm, wm, wm));
        }
      }
    }
#undef PAS_LOAD_T_SLOTS
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

// ---- the ray setup pass -----------------------------------------------------------------------------------
// One block per ray (layer k, mu row j): runs the prologues of the two ray-march kernels and stores their
// results (RaySetupLayout). Same code, same arithmetic: a pass fed from the tables gives the bits it would
// have computed itself.
template <int NC>
__global__ void __launch_bounds__(256, 4)
ray_setup_kernel(const __grid_constant__ PasGeometry g, const float* __restrict__ T, char* __restrict__ setup,
                 int k_begin, int k_stride, int with_rows) {
  constexpr int CP = PAS_CHANNEL_PITCH(NC), WIDTH = 256;
  __shared__ ScatterSample sSample[kSamples];
  __shared__ SunSample sSun[kSamples];
  __shared__ __align__(16) float sTw[kSamples][CP];
  __shared__ __align__(16) int sPerm[WIDTH];
  __shared__ int sCount[WIDTH / 32];
  const int tid = threadIdx.x;
  const int j = blockIdx.x, k = k_begin + blockIdx.y * k_stride;
  const BlockRay ray = block_ray(g, k, j);
  const RaySetupLayout lay = ray_setup_layout(CP);
  char* rec = setup + ((size_t)k * g.sz.mu_n + j) * lay.stride;
  // single scattering: records + path transmittances
  sun_prologue<NC, WIDTH>(g, T, ray, tid, sSun, sTw);
  __syncthreads();
  copy16(rec + lay.sun, sSun, kSamples * 3, tid, WIDTH);
  copy16(rec + lay.tw, sTw, kSamples * CP / 4, tid, WIDTH);
  if (with_rows) {
    // multiple scattering: slot plan + permutation (the path transmittances are the same numbers)
    rows_prologue<NC, false>(g, T, ray, tid, sSample, sTw, sPerm, sCount);
    copy16(rec + lay.plan, sSample, kSamples * 3, tid, WIDTH);
    copy16(rec + lay.perm, sPerm, WIDTH / 4, tid, WIDTH);
  }
}

inline int round_up32(int v) { return (v + 31) / 32 * 32; }

// Rows of up to 256 texels (the reference's 8 x 32) run with 256-thread blocks and a register
// budget that keeps the 128-bit corner loads of a whole texel in flight; wider rows (up to 1024)
// fall back to one big block per row.
template <typename Kern>
cudaError_t prepare(Kern kern, size_t dyn) {
  if (dyn > 32 * 1024) {
    return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
  }
  return cudaSuccess;
}

template <int NC>
cudaError_t launch_multiple_nc(const PasGeometry& g, const PasSpectrum& s, const float* T,
                               const float* dJ, float* dS, FinalTables fin, LayerSet layers,
                               cudaStream_t stream, const void* setup) {
  constexpr int CP = PAS_CHANNEL_PITCH(NC), Q = CP / 4;
  const int width = g.sz.nu_n * g.sz.mu_s_n;
  if (width > 1024) return cudaErrorInvalidValue;
  const int threads = round_up32(width < kSamples ? kSamples : width);
  const size_t dyn = (size_t)2 * Q * (((width + 7) & ~7) + 8 / Q) * sizeof(float4);
  if (layers.count() == 0) return cudaSuccess;
  const dim3 grid(g.sz.mu_n, layers.count());
  cudaError_t e;
  if (Q > 1 && rows_shape_ok(g)) {
    auto kern = multiple_scattering_rows_kernel<NC>;  // the reference's 8 x 32 row
    const size_t dyn_rows = (size_t)2 * Q * rows_pitch<Q>() * sizeof(float4);
    if ((e = prepare(kern, dyn_rows)) != cudaSuccess) return e;
    kern<<<grid, 256, dyn_rows, stream>>>(g, s, T, dJ, dS, fin, layers.begin, layers.stride, static_cast<const char*>(setup));
  } else if (width == 256) {
    auto kern = multiple_scattering_kernel<NC, 256, 3, 256>;
    if ((e = prepare(kern, dyn)) != cudaSuccess) return e;
    kern<<<grid, 256, dyn, stream>>>(g, s, T, dJ, dS, fin, layers.begin, layers.stride);
  } else if (threads <= 256) {
    auto kern = multiple_scattering_kernel<NC, 256, 3, 0>;
    if ((e = prepare(kern, dyn)) != cudaSuccess) return e;
    kern<<<grid, threads, dyn, stream>>>(g, s, T, dJ, dS, fin, layers.begin, layers.stride);
  } else {
    auto kern = multiple_scattering_kernel<NC, 1024, 1, 0>;
    if ((e = prepare(kern, dyn)) != cudaSuccess) return e;
    kern<<<grid, threads, dyn, stream>>>(g, s, T, dJ, dS, fin, layers.begin, layers.stride);
  }
  return cudaGetLastError();
}

template <int NC>
cudaError_t launch_single_nc(const PasGeometry& g, const PasSpectrum& s, const float* T, float* dR,
                             float* dM, FinalTables fin, LayerSet layers,
                             cudaStream_t stream, const void* setup) {
  constexpr int CP = PAS_CHANNEL_PITCH(NC), Q = CP / 4;
  const int width = g.sz.nu_n * g.sz.mu_s_n;
  if (width > 1024) return cudaErrorInvalidValue;
  const int threads = round_up32(width < kSamples ? kSamples : width);
  const size_t dyn = (size_t)2 * Q * (((g.sz.t_w + 7) & ~7) + 8 / Q) * sizeof(float4);
  if (layers.count() == 0) return cudaSuccess;
  const dim3 grid(g.sz.mu_n, layers.count());
  cudaError_t e;
  if (width == 256 && g.sz.t_w == 256) {
    auto kern = single_scattering_kernel<NC, 256, 2, 256, true>;  // 128 registers: the row slots stay in registers
    const size_t dyn_split = (size_t)2 * Q * (2 * (256 / 2 + 4) + (Q > 1 ? 4 / Q : 0)) * sizeof(float4);
    if ((e = prepare(kern, dyn_split)) != cudaSuccess) return e;
    kern<<<grid, 256, dyn_split, stream>>>(g, s, T, dR, dM, fin, layers.begin, layers.stride, static_cast<const char*>(setup));
  } else if (threads <= 256) {
    auto kern = single_scattering_kernel<NC, 256, 3, 0, false>;
    if ((e = prepare(kern, dyn)) != cudaSuccess) return e;
    kern<<<grid, threads, dyn, stream>>>(g, s, T, dR, dM, fin, layers.begin, layers.stride, nullptr);
  } else {
    auto kern = single_scattering_kernel<NC, 1024, 1, 0, false>;
    if ((e = prepare(kern, dyn)) != cudaSuccess) return e;
    kern<<<grid, threads, dyn, stream>>>(g, s, T, dR, dM, fin, layers.begin, layers.stride, nullptr);
  }
  return cudaGetLastError();
}

}  // namespace

cudaError_t launch_multiple_scattering(const PasGeometry& g, const PasSpectrum& s, const float* T,
                                       const float* dJ, float* dS, FinalTables fin, LayerSet layers,
                                       cudaStream_t stream, const void* setup) {
  switch (s.nc) {
#define PAS_CASE(N) \
  case N: return launch_multiple_nc<N>(g, s, T, dJ, dS, fin, layers, stream, setup);
    PAS_CASE(3) PAS_CASE(4) PAS_CASE(8) PAS_CASE(15) PAS_CASE(16)
#undef PAS_CASE
    default: return cudaErrorInvalidValue;
  }
}

cudaError_t launch_single_scattering(const PasGeometry& g, const PasSpectrum& s, const float* T,
                                     float* dR, float* dM, FinalTables fin, LayerSet layers,
                                     cudaStream_t stream, const void* setup) {
  switch (s.nc) {
#define PAS_CASE(N) \
  case N: return launch_single_nc<N>(g, s, T, dR, dM, fin, layers, stream, setup);
    PAS_CASE(3) PAS_CASE(4) PAS_CASE(8) PAS_CASE(15) PAS_CASE(16)
#undef PAS_CASE
    default: return cudaErrorInvalidValue;
  }
}

// Setup tables of the ray-march passes (ray_setup_kernel): 0 bytes when the table sizes do not take the
// register-slot kernels (row width and transmittance width of 256).
size_t ray_setup_bytes(const PasGeometry& g, int nc) {
  if (g.sz.nu_n * g.sz.mu_s_n != 256 || g.sz.t_w != 256) return 0;
  return ray_setup_layout(PAS_CHANNEL_PITCH(nc)).stride * (size_t)g.sz.mu_n * g.sz.r_n;
}

cudaError_t launch_ray_setup(const PasGeometry& g, const PasSpectrum& s, const float* T, void* setup,
                             LayerSet layers, cudaStream_t stream) {
  if (ray_setup_bytes(g, s.nc) == 0 || layers.count() == 0) return cudaSuccess;
  const dim3 grid(g.sz.mu_n, layers.count());
  const int with_rows = PAS_CHANNEL_PITCH(s.nc) > 4 ? 1 : 0;   // the multiple-scattering rows kernel: > 4 channels
  switch (s.nc) {
#define PAS_CASE(N) \
  case N: ray_setup_kernel<N><<<grid, 256, 0, stream>>>(g, T, static_cast<char*>(setup), layers.begin, layers.stride, with_rows); break;
    PAS_CASE(3) PAS_CASE(4) PAS_CASE(8) PAS_CASE(15) PAS_CASE(16)
#undef PAS_CASE
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

}  // namespace pas

// Launch wrappers of the sm_100a kernels (one per pass of atmosphere/model.cc:1048-1215).
// The multi-channel tables (T, dR, dM, dJ, dS) are channel-interleaved fp32 in HBM:
// tab[texel * CP + c], CP = PAS_CHANNEL_PITCH(nc), texel = x + width * (j + mu_n * k) with
// x = i_nu * mu_s_n + i_mu_s, i.e. the reference's texel order with the channels of one texel side
// by side (pas_types.h). The irradiance tables dE are planar [c][j][i]. Every launcher enqueues on
// `stream` and returns the CUDA status of the launch.
#ifndef PAS_B200_CSRC_PAS_KERNELS_H_
#define PAS_B200_CSRC_PAS_KERNELS_H_

#include <cuda_runtime.h>
#include <cuda_fp16.h>

#include "pas_types.h"

namespace pas {

// Output of the passes that also feed the final (possibly luminance-converted) tables.
struct FinalTables {
  void* scattering;        // RGBA interleaved [k][j][x][4], fp32 or fp16
  void* single_mie;        // RGBA interleaved or nullptr (combined textures)
  float* irradiance;       // RGBA interleaved fp32 [j][i][4]
  int half_precision;      // scattering / single_mie stored as __half
  int accumulate;          // 0: first channel group overwrites, 1: adds (GL blend, model.cc:1083)
  void* host_scattering;   // device-visible address of a pinned HOST copy of `scattering`, or nullptr: the
                           // multiple-scattering pass that makes the table final stores every texel there
                           // too, so that the read-back rides on the kernel instead of following it
};

// Multi-GPU: the same table in the memory of the other GPUs of the box (CUDA IPC mappings, NVLink).
// A pass that is given mirrors stores every texel it produces to each of them as well, at the same
// offset, so the exchange of the r-slabs rides on the kernel that computes them.
#define PAS_MAX_PEERS 7
struct PeerTables {
  int n;
  float* tab[PAS_MAX_PEERS];
  int multicast;   // tab[0] is an NVLS multicast address: one store reaches every rank (the local copy too)
};

// The scattering layers a launch works on: begin, begin + stride, ... < end. One GPU: all of them.
// Several GPUs: a contiguous slab per rank, with either exchange (dealing every world-th layer to a
// rank was measured slower; the kernels accept any stride).
struct LayerSet {
  int begin, end, stride;
  int count() const { return end > begin ? (end - begin + stride - 1) / stride : 0; }
};

// T[c][j][i] = transmittance to the top boundary (ComputeTransmittanceToTopAtmosphereBoundaryTexture,
// functions.glsl:454-463). One warp per texel, 501 samples split across lanes, fp64.
// With `rgb` / `rgba`, the same launch also fills the final RGBA32F transmittance texture for the
// three channels of `rgb` (the optical lengths are wavelength independent).
struct RgbExtinction {
  double beta_r[3], beta_m_ext[3], beta_abs[3];
};
cudaError_t launch_transmittance(const PasGeometry& g, const PasSpectrum& s, float* T,
                                 cudaStream_t stream, const PasSpectrum* rgb = nullptr,
                                 float* rgba = nullptr);
// Rows [j_begin, j_end) only, stored to T and to its mirrors (multi-GPU: one band of rows per rank);
// optionally the same rows of the final RGBA table, to `rgba` and its mirrors.
cudaError_t launch_transmittance_rows(const PasGeometry& g, const PasSpectrum& s, float* T,
                                      const PeerTables& mirrors, int j_begin, int j_end,
                                      cudaStream_t stream, const PasSpectrum* rgb = nullptr,
                                      float* rgba = nullptr, const PeerTables* rgba_mirrors = nullptr);
// Packs channels 0..2 of an interleaved transmittance table into the RGBA32F product table.
cudaError_t launch_pack_rgba(const float* table, int n_texels, int nc, float* rgba,
                             cudaStream_t stream);
// Layout conversions for the test hooks / captures, which present tables planar [c][texel].
cudaError_t launch_interleaved_to_planar(const float* src, size_t n_texels, int nc, float* dst,
                                         cudaStream_t stream);
cudaError_t launch_planar_to_interleaved(const float* src, size_t n_texels, int nc, float* dst,
                                         cudaStream_t stream);

// dE[c][j][i] = direct irradiance (functions.glsl:1558-1567); zero-initialises the final E when
// !accumulate (model.cc:139).
cudaError_t launch_direct_irradiance(const PasGeometry& g, const PasSpectrum& s, const float* T,
                                     float* dE, FinalTables fin, cudaStream_t stream);

// Per-(layer, direction) constants of the density pass + ground factor
// G[k][l][c] = T(r_k -> ground along theta_l)[c] * albedo[c] / pi (functions.glsl:1198-1213,1239-1240)
// + per-layer scattering coefficients cR[k][c] = beta_R[c] rho_R(h_k), cM[k][c] = beta_M[c] rho_M(h_k).
cudaError_t launch_density_setup(const PasGeometry& g, const PasSpectrum& s, const float* T,
                                 PasDensityDir* dirs, float* G, float* cR, float* cM,
                                 cudaStream_t stream);

// Single scattering (functions.glsl:933-945) for layers [k_begin, k_end) + fused epilogue
// S.rgb (+)= L.dR, S.a (+)= (L.dM).r, M (+)= L.dM (model.cc:142-157).
cudaError_t launch_single_scattering(const PasGeometry& g, const PasSpectrum& s, const float* T,
                                     float* dR, float* dM, FinalTables fin, LayerSet layers,
                                     cudaStream_t stream, const void* ray_setup = nullptr);

// Per-ray setup tables shared by the single-scattering pass and every multiple-scattering pass of an
// Init (sample records, path transmittances, texel permutation: what their blocks would otherwise
// compute in a latency-bound prologue, once per pass). ray_setup_bytes: size of the table, 0 when the
// table sizes do not take the kernels that use it; the passes then compute their prologues themselves
// (as they do when given ray_setup = nullptr).
size_t ray_setup_bytes(const PasGeometry& g, int nc);
cudaError_t launch_ray_setup(const PasGeometry& g, const PasSpectrum& s, const float* T, void* ray_setup,
                             LayerSet layers, cudaStream_t stream);

// Scattering density of `order` >= 2 (functions.glsl:1348-1367) for layers [k_begin, k_end).
// Reads dR, dM (order 2) or dS (order >= 3) and row 0 of dE. Every texel is also stored to the
// `mirrors` of dJ (none on a single GPU).
cudaError_t launch_scattering_density(const PasGeometry& g, const PasSpectrum& s,
                                      const PasDensityDir* dirs, const float* G, const float* cR,
                                      const float* cM, const float* dR, const float* dM,
                                      const float* dS, const float* dE, int order, float* dJ,
                                      const PeerTables& mirrors, LayerSet layers,
                                      cudaStream_t stream);

// Indirect irradiance from radiance of `order` (1: dR/dM with phase functions, else dS)
// (functions.glsl:1573-1586) for rows [j_begin, j_end) + fused E += L.dE (model.cc:176-190)
// when fin.irradiance != nullptr. Only source layers [k_begin, k_end) contribute (the r-slab
// owned by this rank; the partial sums are all-reduced by the caller).
cudaError_t launch_indirect_irradiance(const PasGeometry& g, const PasSpectrum& s, const float* dR,
                                       const float* dM, const float* dS, int order, float* dE,
                                       FinalTables fin, int j_begin, int j_end, LayerSet layers,
                                       cudaStream_t stream);

// Multiple scattering (functions.glsl:1369-1383) for layers [k_begin, k_end) + fused
// S.rgb += L.dS / RayleighPhaseFunction(nu) (model.cc:192-208).
cudaError_t launch_multiple_scattering(const PasGeometry& g, const PasSpectrum& s, const float* T,
                                       const float* dJ, float* dS, FinalTables fin, LayerSet layers,
                                       cudaStream_t stream, const void* ray_setup = nullptr);

// ---- multi-GPU exchange over peer memory (peer_exchange.cu) ---------------------------------------
// Flag array of a rank: PAS_FLAG_CHANNELS independent barrier sequences ("channels": main stream, side
// stream) of PAS_FLAG_WORDS words each. Word p < 8 of a channel: the last epoch rank p has signalled
// (written by rank p); word PAS_FLAG_EPOCH: the rank's own epoch counter of the channel, advanced by
// its barrier kernels themselves (no host-side epoch: the same enqueued sequence can be replayed);
// word PAS_FLAG_POISON of channel 0: set by any rank whose barrier timed out -- every later barrier of
// every rank then fails at once instead of running on with stale data.
#define PAS_FLAG_CHANNELS 2
#define PAS_FLAG_WORDS 16
#define PAS_FLAG_EPOCH 8
#define PAS_FLAG_POISON 15
struct PeerFlags {                 // flags[r]: the channel's words in the flag array of rank r
  int rank, world;
  unsigned* flags[PAS_MAX_PEERS + 1];
  unsigned* poison[PAS_MAX_PEERS + 1];   // the poison word of rank r
  int* error;                      // mapped host memory: 1 = this rank timed out, 2 = poisoned by a peer
};
struct PeerTargets {               // destinations of a push: the same buffer on every rank
  int n;
  void* dst[PAS_MAX_PEERS + 1];
};
// Next cross-GPU barrier of the channel (epochs advance by one per barrier, in step on every rank).
cudaError_t launch_peer_barrier(const PeerFlags& f, cudaStream_t stream);
// Copies `chunks` byte ranges [offset + i * stride, offset + i * stride + bytes) of src to the same
// ranges of every target.
cudaError_t launch_peer_push(const void* src, size_t bytes, size_t offset_bytes, const PeerTargets& t,
                             cudaStream_t stream, int chunks = 1, size_t stride_bytes = 0);
// out[i] = sum over ranks r (in rank order) of parts[r * stride + i].
cudaError_t launch_sum_partials(const float* parts, int world, size_t stride, int n, float* out,
                                cudaStream_t stream);

// ---- render-time use of the tables (kernel_render.cuh: device functions; kernel_render.cu: batches) ---
struct RenderTables {
  const float4* transmittance;  // RGBA32F [t_h][t_w]
  const void* scattering;       // RGBA [r][mu][nu * mu_s], fp32 or fp16
  const void* single_mie;       // RGBA or nullptr (combined textures: alpha of `scattering`)
  const float4* irradiance;     // RGBA32F [e_h][e_w]
  int half_precision;
};
struct RenderConstants {         // the ATMOSPHERE constants the render functions read, at 680/550/440 nm
  double solar[3], rayleigh[3], mie_sca[3];
  double sky_k[3], sun_k[3];     // SKY / SUN_SPECTRAL_RADIANCE_TO_LUMINANCE, or 1 (radiance mode)
};
// GetSkyRadiance (to_point = false: target = view ray) / GetSkyRadianceToPoint (target = point) for
// n queries; device pointers, vectors [n][3]; shadow_length and transmittance may be nullptr.
cudaError_t launch_sky_radiance(const PasGeometry& g, const RenderTables& t, const RenderConstants& c,
                                size_t n, bool to_point, const double* camera, const double* target,
                                const double* shadow_length, const double* sun_direction, float* radiance,
                                float* transmittance, cudaStream_t stream);
// GetSunAndSkyIrradiance for n queries.
cudaError_t launch_sun_and_sky_irradiance(const PasGeometry& g, const RenderTables& t,
                                          const RenderConstants& c, size_t n, const double* point,
                                          const double* normal, const double* sun_direction,
                                          float* sun_irradiance, float* sky_irradiance, cudaStream_t stream);

// Channel-group sizes the templated kernels are instantiated for.
bool channel_count_supported(int nc);

}  // namespace pas

#endif  // PAS_B200_CSRC_PAS_KERNELS_H_

// Dispatch of the scattering density pass (kernel_density.cuh holds the kernel; its instantiations are
// compiled in kernel_density_cp*.cu, one translation unit per (channel pitch, single / multi GPU) so that
// they build in parallel).
#include "kernel_density.cuh"

namespace pas {
namespace density {
extern template cudaError_t launch_cp<4, false>(const PasGeometry& g, int nc, const PasDensityDir* dirs, const float* G,
                                const float* cR, const float* cM, const float* dR, const float* dM,
                                const float* dS, const float* dE, int order, float* dJ,
                                const PeerTables& mirrors, LayerSet layers, cudaStream_t stream);
extern template cudaError_t launch_cp<4, true>(const PasGeometry& g, int nc, const PasDensityDir* dirs, const float* G,
                                const float* cR, const float* cM, const float* dR, const float* dM,
                                const float* dS, const float* dE, int order, float* dJ,
                                const PeerTables& mirrors, LayerSet layers, cudaStream_t stream);
extern template cudaError_t launch_cp<8, false>(const PasGeometry& g, int nc, const PasDensityDir* dirs, const float* G,
                                const float* cR, const float* cM, const float* dR, const float* dM,
                                const float* dS, const float* dE, int order, float* dJ,
                                const PeerTables& mirrors, LayerSet layers, cudaStream_t stream);
extern template cudaError_t launch_cp<8, true>(const PasGeometry& g, int nc, const PasDensityDir* dirs, const float* G,
                                const float* cR, const float* cM, const float* dR, const float* dM,
                                const float* dS, const float* dE, int order, float* dJ,
                                const PeerTables& mirrors, LayerSet layers, cudaStream_t stream);
extern template cudaError_t launch_cp<16, false>(const PasGeometry& g, int nc, const PasDensityDir* dirs, const float* G,
                                const float* cR, const float* cM, const float* dR, const float* dM,
                                const float* dS, const float* dE, int order, float* dJ,
                                const PeerTables& mirrors, LayerSet layers, cudaStream_t stream);
extern template cudaError_t launch_cp<16, true>(const PasGeometry& g, int nc, const PasDensityDir* dirs, const float* G,
                                const float* cR, const float* cM, const float* dR, const float* dM,
                                const float* dS, const float* dE, int order, float* dJ,
                                const PeerTables& mirrors, LayerSet layers, cudaStream_t stream);
}  // namespace density

cudaError_t launch_scattering_density(const PasGeometry& g, const PasSpectrum& s,
                                      const PasDensityDir* dirs, const float* G, const float* cR,
                                      const float* cM, const float* dR, const float* dM,
                                      const float* dS, const float* dE, int order, float* dJ,
                                      const PeerTables& mirrors, LayerSet layers,
                                      cudaStream_t stream) {
  if (g.sz.nu_n < 2 || g.sz.nu_n > PAS_MAX_NU) return cudaErrorInvalidValue;
  if (!channel_count_supported(s.nc)) return cudaErrorInvalidValue;
  const bool mirror = mirrors.n > 0;
  switch (PAS_CHANNEL_PITCH(s.nc)) {
#define PAS_CASE(CP)                                                                                   \
  case CP:                                                                                             \
    return mirror ? density::launch_cp<CP, true>(g, s.nc, dirs, G, cR, cM, dR, dM, dS, dE, order, dJ, mirrors, layers, stream)                                        \
                  : density::launch_cp<CP, false>(g, s.nc, dirs, G, cR, cM, dR, dM, dS, dE, order, dJ, mirrors, layers, stream);
    PAS_CASE(4) PAS_CASE(8) PAS_CASE(16)
#undef PAS_CASE
    default: return cudaErrorInvalidValue;
  }
}

bool channel_count_supported(int nc) {
  return nc == 3 || nc == 4 || nc == 8 || nc == 15 || nc == 16;
}

}  // namespace pas

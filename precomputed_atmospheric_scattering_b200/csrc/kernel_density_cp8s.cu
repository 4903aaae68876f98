// Scattering density kernel, channel pitch 8, single GPU instantiations.
#include "kernel_density.cuh"

namespace pas {
namespace density {
template cudaError_t launch_cp<8, false>(const PasGeometry& g, int nc, const PasDensityDir* dirs, const float* G,
                                const float* cR, const float* cM, const float* dR, const float* dM,
                                const float* dS, const float* dE, int order, float* dJ,
                                const PeerTables& mirrors, LayerSet layers, cudaStream_t stream);
}  // namespace density
}  // namespace pas

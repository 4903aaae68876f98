"""B200-native LUT precomputation for ebruneton/precomputed_atmospheric_scattering.

The product is the C-ABI library ``libpas_b200.so`` (include/pas_b200.h, sources under csrc/): the
hand-written sm_100a kernels and the host schedule that replace ``atmosphere::Model::Init``. This
package is the thin Python mirror of the reference's ``atmosphere::Model`` over that ABI, plus the
atmosphere definitions of the bench configs.
"""
from .atmospheres import (AtmosphereSpec, ChannelParams, DensityProfileLayer, channel_params, earth,
                          model_test_earth, precomputed_wavelengths, small_planet)
from . import ensemble
from .model import (LIB_PATH, Model, PasError, TEXTURE_IRRADIANCE, TEXTURE_SCATTERING,
                    TEXTURE_SINGLE_MIE, TEXTURE_TRANSMITTANCE, convert_spectrum_to_linear_srgb,
                    load_library, measure_device_peaks, nccl_unique_id, release_cached_memory,
                    shader_source, spectral_channels, world_is_cached)

__all__ = ["AtmosphereSpec", "ChannelParams", "DensityProfileLayer", "Model", "PasError", "LIB_PATH",
           "channel_params", "earth", "small_planet", "model_test_earth", "ensemble", "precomputed_wavelengths", "load_library",
           "nccl_unique_id", "shader_source", "spectral_channels", "measure_device_peaks", "release_cached_memory", "world_is_cached", "convert_spectrum_to_linear_srgb", "TEXTURE_TRANSMITTANCE",
           "TEXTURE_SCATTERING", "TEXTURE_IRRADIANCE", "TEXTURE_SINGLE_MIE"]

"""Atmosphere definitions used by the bench configs and the parity tests.

The Earth atmosphere below restates the *published physical input data* the reference demo
feeds its model (reference: atmosphere/demo/demo.cc:194-276 and
atmosphere/reference/model_test.cc:222-308): the ASTM G-173 extraterrestrial solar spectrum
binned to 10 nm, the Serdyuchenko/Bremen 233 K ozone cross-sections binned to 10 nm, the
Rayleigh/Mie constants and the two-layer ozone profile. They are inputs of the hot path, not
part of it.

Everything here is plain host-side data in SI units, exactly what the reference's
``atmosphere::Model`` constructor takes (atmosphere/model.h:182-281).
"""
from __future__ import annotations

import dataclasses
import math
from typing import List, Sequence

import numpy as np

# atmosphere/model.h:306-308
LAMBDA_R, LAMBDA_G, LAMBDA_B = 680.0, 550.0, 440.0
LAMBDA_MIN, LAMBDA_MAX = 360.0, 830.0

# ASTM G-173 ETR, averaged per 10 nm bin from 360 nm, W/m^2/nm (demo.cc:196-203).
SOLAR_IRRADIANCE = [
    1.11776, 1.14259, 1.01249, 1.14716, 1.72765, 1.73054, 1.6887, 1.61253,
    1.91198, 2.03474, 2.02042, 2.02212, 1.93377, 1.95809, 1.91686, 1.8298,
    1.8685, 1.8931, 1.85149, 1.8504, 1.8341, 1.8345, 1.8147, 1.78158, 1.7533,
    1.6965, 1.68194, 1.64654, 1.6048, 1.52143, 1.55622, 1.5113, 1.474, 1.4482,
    1.41018, 1.36775, 1.34188, 1.31429, 1.28303, 1.26758, 1.2367, 1.2082,
    1.18737, 1.14683, 1.12362, 1.1058, 1.07124, 1.04992,
]
# Ozone absorption cross-section at 233 K, averaged per 10 nm bin, m^2 (demo.cc:208-222).
OZONE_CROSS_SECTION = [
    1.18e-27, 2.182e-28, 2.818e-28, 6.636e-28, 1.527e-27, 2.763e-27, 5.52e-27,
    8.451e-27, 1.582e-26, 2.316e-26, 3.669e-26, 4.924e-26, 7.752e-26, 9.016e-26,
    1.48e-25, 1.602e-25, 2.139e-25, 2.755e-25, 3.091e-25, 3.5e-25, 4.266e-25,
    4.672e-25, 4.398e-25, 4.701e-25, 5.019e-25, 4.305e-25, 3.74e-25, 3.215e-25,
    2.662e-25, 2.238e-25, 1.852e-25, 1.473e-25, 1.209e-25, 9.423e-26, 7.455e-26,
    6.566e-26, 5.105e-26, 4.15e-26, 4.228e-26, 3.237e-26, 2.451e-26, 2.801e-26,
    2.534e-26, 1.624e-26, 1.465e-26, 2.078e-26, 1.383e-26, 7.105e-27,
]
DOBSON_UNIT = 2.687e20  # molecules / m^2


@dataclasses.dataclass
class DensityProfileLayer:
    """atmosphere/model.h:165-178. density = exp_term*exp(exp_scale*h) + linear_term*h + constant_term."""
    width: float = 0.0
    exp_term: float = 0.0
    exp_scale: float = 0.0
    linear_term: float = 0.0
    constant_term: float = 0.0

    def astuple(self):
        return (self.width, self.exp_term, self.exp_scale, self.linear_term, self.constant_term)


@dataclasses.dataclass
class AtmosphereSpec:
    """The 19 constructor arguments of atmosphere::Model (atmosphere/model.h:182-281), SI units."""
    wavelengths: List[float]
    solar_irradiance: List[float]
    sun_angular_radius: float
    bottom_radius: float
    top_radius: float
    rayleigh_density: List[DensityProfileLayer]
    rayleigh_scattering: List[float]
    mie_density: List[DensityProfileLayer]
    mie_scattering: List[float]
    mie_extinction: List[float]
    mie_phase_function_g: float
    absorption_density: List[DensityProfileLayer]
    absorption_extinction: List[float]
    ground_albedo: List[float]
    max_sun_zenith_angle: float
    length_unit_in_meters: float = 1000.0
    num_precomputed_wavelengths: int = 3
    combine_scattering_textures: bool = True
    half_precision: bool = False


def earth(num_precomputed_wavelengths: int = 3, *, half_precision: bool = False,
          combine_scattering_textures: bool = True, use_ozone: bool = True,
          mie_scale_height: float = 1200.0, ozone_dobson: float = 300.0,
          ground_albedo: float = 0.1, max_sun_zenith_deg: float | None = None,
          sun_angular_radius: float = 0.00935 / 2.0) -> AtmosphereSpec:
    """Earth atmosphere with the demo's parameters (demo.cc:188-284).

    ``max_sun_zenith_deg`` defaults to the demo's rule: 102 deg with half precision tables, else
    120 deg (demo.cc:236-237). The demo's own pi constant (3.1415926, demo.cc:69) is kept so the
    angle handed to the model is bit-identical to the demo's.
    """
    k_pi = 3.1415926
    if max_sun_zenith_deg is None:
        max_sun_zenith_deg = 102.0 if half_precision else 120.0
    k_rayleigh = 1.24062e-6
    rayleigh_scale_height = 8000.0
    mie_angstrom_alpha, mie_angstrom_beta = 0.0, 5.328e-3
    mie_ssa, mie_g = 0.9, 0.8
    max_ozone_number_density = ozone_dobson * DOBSON_UNIT / 15000.0
    wl, sol, ray, mie_s, mie_e, absorb, alb = [], [], [], [], [], [], []
    for idx, l in enumerate(range(360, 831, 10)):
        lam = l * 1e-3
        mie = mie_angstrom_beta / mie_scale_height * lam ** (-mie_angstrom_alpha)
        wl.append(float(l))
        sol.append(SOLAR_IRRADIANCE[idx])
        ray.append(k_rayleigh * lam ** -4)
        mie_s.append(mie * mie_ssa)
        mie_e.append(mie)
        absorb.append(max_ozone_number_density * OZONE_CROSS_SECTION[idx] if use_ozone else 0.0)
        alb.append(ground_albedo)
    return AtmosphereSpec(
        wavelengths=wl, solar_irradiance=sol, sun_angular_radius=sun_angular_radius,
        bottom_radius=6360000.0, top_radius=6420000.0,
        rayleigh_density=[DensityProfileLayer(0.0, 1.0, -1.0 / rayleigh_scale_height, 0.0, 0.0)],
        rayleigh_scattering=ray,
        mie_density=[DensityProfileLayer(0.0, 1.0, -1.0 / mie_scale_height, 0.0, 0.0)],
        mie_scattering=mie_s, mie_extinction=mie_e, mie_phase_function_g=mie_g,
        absorption_density=[DensityProfileLayer(25000.0, 0.0, 0.0, 1.0 / 15000.0, -2.0 / 3.0),
                            DensityProfileLayer(0.0, 0.0, 0.0, -1.0 / 15000.0, 8.0 / 3.0)],
        absorption_extinction=absorb, ground_albedo=alb,
        max_sun_zenith_angle=max_sun_zenith_deg / 180.0 * k_pi,
        length_unit_in_meters=1000.0,
        num_precomputed_wavelengths=num_precomputed_wavelengths,
        combine_scattering_textures=combine_scattering_textures,
        half_precision=half_precision)


def model_test_earth(num_precomputed_wavelengths: int = 3, *, combine_scattering_textures: bool = True,
                     half_precision: bool = True, ground_albedo: float = 0.1) -> AtmosphereSpec:
    """The atmosphere of the reference's integration test (reference/model_test.cc:222-308): the
    demo's Earth with a sun of angular radius 0.2678 deg and mu_s_min = cos(102 deg)."""
    spec = earth(num_precomputed_wavelengths, half_precision=half_precision,
                 combine_scattering_textures=combine_scattering_textures, ground_albedo=ground_albedo,
                 sun_angular_radius=0.2678 * math.pi / 180.0)
    spec.max_sun_zenith_angle = 102.0 * math.pi / 180.0
    return spec


def small_planet() -> AtmosphereSpec:
    """The synthetic planet of the reference's unit tests (reference/functions_test.cc:51-61):
    radii 1000/1500 km, Rayleigh/Mie scale heights 60/30 km. Extended here with plausible spectra so
    that every pass has non-trivial input; used with reduced table sizes in the fast parity tests."""
    spec = earth(3)
    spec.bottom_radius, spec.top_radius = 1000e3, 1500e3
    spec.rayleigh_density = [DensityProfileLayer(0.0, 1.0, -1.0 / 60e3, 0.0, 0.0)]
    spec.mie_density = [DensityProfileLayer(0.0, 1.0, -1.0 / 30e3, 0.0, 0.0)]
    spec.absorption_density = [DensityProfileLayer(250e3, 0.0, 0.0, 1.0 / 150e3, -2.0 / 3.0),
                               DensityProfileLayer(0.0, 0.0, 0.0, -1.0 / 150e3, 8.0 / 3.0)]
    # keep optical depths comparable to Earth's: coefficients scale with 1/scale height
    spec.rayleigh_scattering = [v * 8.0 / 60.0 for v in spec.rayleigh_scattering]
    spec.mie_scattering = [v * 1.2 / 30.0 for v in spec.mie_scattering]
    spec.mie_extinction = [v * 1.2 / 30.0 for v in spec.mie_extinction]
    spec.absorption_extinction = [v * 0.1 for v in spec.absorption_extinction]
    spec.sun_angular_radius = 0.02
    return spec


def interpolate(wavelengths: Sequence[float], values: Sequence[float], wavelength: float) -> float:
    """Piecewise-linear spectrum lookup, clamped at both ends (atmosphere/model.cc:535-552)."""
    if wavelength < wavelengths[0]:
        return values[0]
    for i in range(len(wavelengths) - 1):
        if wavelength < wavelengths[i + 1]:
            u = (wavelength - wavelengths[i]) / (wavelengths[i + 1] - wavelengths[i])
            return values[i] * (1.0 - u) + values[i + 1] * u
    return values[-1]


def precomputed_wavelengths(num_precomputed_wavelengths: int) -> List[float]:
    """Wavelengths the reference precomputes (atmosphere/model.cc:907-924): the RGB triple for
    n <= 3, else 3*ceil(n/3) band centres over [360, 830] nm."""
    if num_precomputed_wavelengths <= 3:
        return [LAMBDA_R, LAMBDA_G, LAMBDA_B]
    iters = (num_precomputed_wavelengths + 2) // 3
    dl = (LAMBDA_MAX - LAMBDA_MIN) / (3 * iters)
    return [LAMBDA_MIN + (j + 0.5) * dl for j in range(3 * iters)]


@dataclasses.dataclass
class ChannelParams:
    """Per-channel physical parameters in *length units* (km for the demo), i.e. the numbers the
    reference bakes into its GLSL header (atmosphere/model.cc:718-734) but without the 6-decimal
    ``std::to_string`` truncation (our tables follow the CPU oracle)."""
    lambdas: np.ndarray
    solar_irradiance: np.ndarray
    rayleigh_scattering: np.ndarray
    mie_scattering: np.ndarray
    mie_extinction: np.ndarray
    absorption_extinction: np.ndarray
    ground_albedo: np.ndarray
    sun_angular_radius: float
    bottom_radius: float
    top_radius: float
    mie_phase_function_g: float
    mu_s_min: float
    profiles: np.ndarray  # [3 profiles][2 layers][5]: rayleigh, mie, absorption

    @property
    def num_channels(self) -> int:
        return len(self.lambdas)


def _pad_layers(layers: Sequence[DensityProfileLayer], unit: float) -> np.ndarray:
    """Missing layers are padded at the FRONT with zero layers (atmosphere/model.cc:653-666);
    widths are divided and inverse lengths multiplied by the length unit (model.cc:641-650)."""
    ls = list(layers)
    if len(ls) > 2:
        raise ValueError("at most 2 density layers (atmosphere/model.h:206-207)")
    while len(ls) < 2:
        ls.insert(0, DensityProfileLayer())
    return np.array([[l.width / unit, l.exp_term, l.exp_scale * unit, l.linear_term * unit,
                      l.constant_term] for l in ls], dtype=np.float64)


def channel_params(spec: AtmosphereSpec, lambdas: Sequence[float]) -> ChannelParams:
    u = spec.length_unit_in_meters
    f = lambda v, scale: np.array([interpolate(spec.wavelengths, v, l) * scale for l in lambdas])
    return ChannelParams(
        lambdas=np.array(lambdas, dtype=np.float64),
        solar_irradiance=f(spec.solar_irradiance, 1.0),
        rayleigh_scattering=f(spec.rayleigh_scattering, u),
        mie_scattering=f(spec.mie_scattering, u),
        mie_extinction=f(spec.mie_extinction, u),
        absorption_extinction=f(spec.absorption_extinction, u),
        ground_albedo=f(spec.ground_albedo, 1.0),
        sun_angular_radius=spec.sun_angular_radius,
        bottom_radius=spec.bottom_radius / u, top_radius=spec.top_radius / u,
        mie_phase_function_g=spec.mie_phase_function_g,
        mu_s_min=math.cos(spec.max_sun_zenith_angle),
        profiles=np.stack([_pad_layers(spec.rayleigh_density, u), _pad_layers(spec.mie_density, u),
                           _pad_layers(spec.absorption_density, u)]))

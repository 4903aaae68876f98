"""Python mirror of the reference's ``atmosphere::Model`` (atmosphere/model.h:180-337) over the C ABI
of libpas_b200.so (include/pas_b200.h). Same constructor arguments, same ``Init``; the tables come
back as numpy arrays instead of GL texture names. This module is plumbing only: every table is
computed by the CUDA kernels behind ``pas_model_init`` and there is no CPU fallback -- if the
library is missing or no GPU is present the constructor raises.
"""
from __future__ import annotations

import array
import ctypes
import os
from typing import Dict, List, Optional, Sequence

import numpy as np

from .atmospheres import AtmosphereSpec, DensityProfileLayer

_HERE = os.path.dirname(os.path.abspath(__file__))
# PAS_B200_LIB points at an alternative build of the same library (A/B runs of kernel variants)
LIB_PATH = os.environ.get("PAS_B200_LIB") or os.path.join(_HERE, "libpas_b200.so")

TEXTURE_TRANSMITTANCE, TEXTURE_SCATTERING, TEXTURE_IRRADIANCE, TEXTURE_SINGLE_MIE = 0, 1, 2, 3
PHASES = {"transmittance": 0, "direct_irradiance": 1, "single_scattering": 2,
          "scattering_density": 3, "indirect_irradiance": 4, "multiple_scattering": 5,
          "density_setup": 6}


class PasError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"pas_b200 status {status}: {message}")
        self.status = status


class _Layer(ctypes.Structure):
    _fields_ = [(n, ctypes.c_double) for n in
                ("width", "exp_term", "exp_scale", "linear_term", "constant_term")]


class _Sizes(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int) for n in
                ("transmittance_width", "transmittance_height", "scattering_r", "scattering_mu",
                 "scattering_mu_s", "scattering_nu", "irradiance_width", "irradiance_height")]


_DP = ctypes.POINTER(ctypes.c_double)
_LP = ctypes.POINTER(_Layer)


class _Params(ctypes.Structure):
    _fields_ = [
        ("num_wavelengths", ctypes.c_size_t), ("wavelengths", _DP), ("solar_irradiance", _DP),
        ("sun_angular_radius", ctypes.c_double), ("bottom_radius", ctypes.c_double),
        ("top_radius", ctypes.c_double),
        ("num_rayleigh_layers", ctypes.c_size_t), ("rayleigh_density", _LP),
        ("rayleigh_scattering", _DP),
        ("num_mie_layers", ctypes.c_size_t), ("mie_density", _LP), ("mie_scattering", _DP),
        ("mie_extinction", _DP), ("mie_phase_function_g", ctypes.c_double),
        ("num_absorption_layers", ctypes.c_size_t), ("absorption_density", _LP),
        ("absorption_extinction", _DP), ("ground_albedo", _DP),
        ("max_sun_zenith_angle", ctypes.c_double), ("length_unit_in_meters", ctypes.c_double),
        ("num_precomputed_wavelengths", ctypes.c_uint), ("combine_scattering_textures", ctypes.c_int),
        ("half_precision", ctypes.c_int), ("sizes", _Sizes), ("device", ctypes.c_int)]


class _TextureInfo(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int) for n in
                ("width", "height", "depth", "channels", "bytes_per_channel", "present")]


_lib = None


IPC_EXPORT_BYTES = 464  # PAS_IPC_EXPORT_BYTES


def load_library() -> ctypes.CDLL:
    """Loads the in-tree C-ABI library. Fails loudly if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FileNotFoundError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
                f"g.build()'` (nvcc, sm_100a). There is no CPU fallback.")
        lib = ctypes.CDLL(LIB_PATH)
        lib.pas_last_error.restype = ctypes.c_char_p
        lib.pas_model_create.argtypes = [ctypes.POINTER(_Params), ctypes.POINTER(ctypes.c_void_p)]
        lib.pas_model_destroy.argtypes = [ctypes.c_void_p]
        lib.pas_model_init.argtypes = [ctypes.c_void_p, ctypes.c_uint]
        lib.pas_model_init_async.argtypes = [ctypes.c_void_p, ctypes.c_uint]
        lib.pas_model_wait.argtypes = [ctypes.c_void_p]
        lib.pas_model_set_host_outputs.argtypes = [ctypes.c_void_p] + [ctypes.c_void_p] * 4
        lib.pas_model_set_host_output_mode.argtypes = [ctypes.c_void_p, ctypes.c_int]
        lib.pas_model_texture_info.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(_TextureInfo)]
        lib.pas_model_texture_device_ptr.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]
        lib.pas_model_read_texture.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t]
        lib.pas_model_save_dat.argtypes = [ctypes.c_void_p, ctypes.c_char_p]
        lib.pas_model_shader_source.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p, ctypes.POINTER(ctypes.c_size_t)]
        lib.pas_shader_source.argtypes = [ctypes.POINTER(_Params), ctypes.c_char_p, ctypes.c_void_p, ctypes.POINTER(ctypes.c_size_t)]
        lib.pas_model_save_webgl.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p]
        lib.pas_model_luminance_factors.argtypes = [ctypes.c_void_p, _DP]
        lib.pas_convert_spectrum_to_linear_srgb.argtypes = [ctypes.c_size_t, _DP, _DP, _DP, _DP, _DP]
        lib.pas_model_channels.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int), _DP]
        lib.pas_model_luminance_matrix.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_float)]
        lib.pas_spectral_channels.argtypes = [ctypes.c_uint, ctypes.POINTER(ctypes.c_int), _DP, ctypes.POINTER(ctypes.c_float)]
        lib.pas_model_set_capture.argtypes = [ctypes.c_void_p, ctypes.c_int]
        lib.pas_model_read_intermediate.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p, ctypes.POINTER(ctypes.c_size_t)]
        lib.pas_model_write_intermediate.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p, ctypes.c_size_t]
        lib.pas_model_run_phase.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
        lib.pas_model_last_timings.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(ctypes.c_float)]
        lib.pas_model_last_launch_count.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int)]
        lib.pas_nccl_unique_id.argtypes = [ctypes.c_void_p]
        lib.pas_model_attach_world.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
        lib.pas_world_is_cached.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int]
        lib.pas_model_ipc_export.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.POINTER(ctypes.c_size_t)]
        lib.pas_model_attach_peers.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]
        lib.pas_model_exchange_bytes.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_size_t)]
        lib.pas_model_attach_symmetric.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                                   ctypes.POINTER(ctypes.c_void_p), ctypes.c_void_p, ctypes.c_size_t]
        lib.pas_release_cached_memory.restype = None
        _FP = ctypes.POINTER(ctypes.c_float)
        lib.pas_model_get_solar_radiance.argtypes = [ctypes.c_void_p, ctypes.c_int, _DP]
        lib.pas_model_get_sky_radiance.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t,
                                                   ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                                   ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        lib.pas_model_get_sky_radiance_to_point.argtypes = lib.pas_model_get_sky_radiance.argtypes
        lib.pas_model_get_sun_and_sky_irradiance.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t,
                                                             ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                                             ctypes.c_void_p, ctypes.c_void_p]
        lib.pas_model_render_context.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                                 ctypes.POINTER(ctypes.c_size_t)]
        lib.pas_model_last_render_ms.argtypes = [ctypes.c_void_p, _FP]
        _lib = lib
    return _lib


def _check(status: int):
    if status != 0:
        raise PasError(status, load_library().pas_last_error().decode("utf-8", "replace"))


def _darr(v: Sequence[float]):
    # (array('d') + from_buffer: a few microseconds; numpy's .ctypes accessor costs 10x that, and a model is
    # created per settings change -- the constructor is part of the end-to-end time)
    a = array.array("d", v)
    return a, ctypes.cast((ctypes.c_double * len(a)).from_buffer(a), _DP)


def _layers(layers: Sequence[DensityProfileLayer]):
    flat = array.array("d", [x for l in layers for x in l.astuple()] or [0.0] * 5)
    return (_Layer * max(len(layers), 1)).from_buffer(flat), flat


def nccl_unique_id() -> bytes:
    buf = ctypes.create_string_buffer(128)
    _check(load_library().pas_nccl_unique_id(buf))
    return buf.raw


def convert_spectrum_to_linear_srgb(wavelengths: Sequence[float], spectrum: Sequence[float]):
    """atmosphere::Model::ConvertSpectrumToLinearSrgb (atmosphere/model.cc:1020-1040)."""
    w, wp = _darr(wavelengths)
    s, sp = _darr(spectrum)
    r, g, b = ctypes.c_double(), ctypes.c_double(), ctypes.c_double()
    _check(load_library().pas_convert_spectrum_to_linear_srgb(
        len(w), wp, sp, ctypes.byref(r), ctypes.byref(g), ctypes.byref(b)))
    return r.value, g.value, b.value


def world_is_cached(device: int, rank: int, world_size: int) -> bool:
    """True once this process holds an NCCL communicator for (device, rank, world_size)."""
    return bool(load_library().pas_world_is_cached(device, rank, world_size))


def release_cached_memory() -> None:
    load_library().pas_release_cached_memory()


def measure_device_peaks(device: int = 0):
    """{'fp32_tflops', 'mufu_gops', 'sm_count'} measured on the device (FMA / MUFU microbenchmarks)."""
    lib = load_library()
    lib.pas_measure_device_peaks.argtypes = [ctypes.c_int, _DP, _DP, ctypes.POINTER(ctypes.c_int)]
    f, u, n = ctypes.c_double(), ctypes.c_double(), ctypes.c_int()
    _check(lib.pas_measure_device_peaks(device, ctypes.byref(f), ctypes.byref(u), ctypes.byref(n)))
    return {"fp32_tflops": f.value, "mufu_gops": u.value, "sm_count": n.value}


def spectral_channels(num_precomputed_wavelengths: int):
    """(lambdas[C], luminance_from_radiance[3, C]) of atmosphere/model.cc:907-943; needs no GPU."""
    lib = load_library()
    n = ctypes.c_int()
    _check(lib.pas_spectral_channels(int(num_precomputed_wavelengths), ctypes.byref(n), None, None))
    lam = np.empty(n.value, dtype=np.float64)
    lum = np.empty((3, n.value), dtype=np.float32)
    _check(lib.pas_spectral_channels(int(num_precomputed_wavelengths), ctypes.byref(n),
                                     lam.ctypes.data_as(_DP),
                                     lum.ctypes.data_as(ctypes.POINTER(ctypes.c_float))))
    return lam, lum


def _make_params(wavelengths, solar_irradiance, sun_angular_radius, bottom_radius, top_radius,
                 rayleigh_density, rayleigh_scattering, mie_density, mie_scattering, mie_extinction,
                 mie_phase_function_g, absorption_density, absorption_extinction, ground_albedo,
                 max_sun_zenith_angle, length_unit_in_meters, num_precomputed_wavelengths,
                 combine_scattering_textures, half_precision, sizes=None, device=None):
    """pas_model_params from the 19 constructor arguments; returns (params, objects to keep alive)."""
    keep = []
    p = _Params()
    p.num_wavelengths = len(wavelengths)
    for name, v in (("wavelengths", wavelengths), ("solar_irradiance", solar_irradiance),
                    ("rayleigh_scattering", rayleigh_scattering), ("mie_scattering", mie_scattering),
                    ("mie_extinction", mie_extinction),
                    ("absorption_extinction", absorption_extinction), ("ground_albedo", ground_albedo)):
        if len(v) != len(wavelengths):
            raise ValueError(f"{name} must have one value per wavelength")  # model.cc:539
        a, ptr = _darr(v)
        keep.append(a)
        setattr(p, name, ptr)
    p.sun_angular_radius, p.bottom_radius, p.top_radius = sun_angular_radius, bottom_radius, top_radius
    for name, layers in (("rayleigh", rayleigh_density), ("mie", mie_density),
                         ("absorption", absorption_density)):
        arr, flat = _layers(layers)
        keep.extend((arr, flat))
        setattr(p, f"num_{name}_layers", len(layers))
        setattr(p, f"{name}_density", ctypes.cast(arr, _LP))
    p.mie_phase_function_g = mie_phase_function_g
    p.max_sun_zenith_angle, p.length_unit_in_meters = max_sun_zenith_angle, length_unit_in_meters
    p.num_precomputed_wavelengths = int(num_precomputed_wavelengths)
    p.combine_scattering_textures = int(bool(combine_scattering_textures))
    p.half_precision = int(bool(half_precision))
    for k, v in (sizes or {}).items():
        setattr(p.sizes, k, int(v))
    p.device = 0 if device is None else device + 1
    return p, keep


def shader_source(spec: AtmosphereSpec, glsl_directory: str, sizes: Optional[Dict[str, int]] = None) -> str:
    """The GLSL source atmosphere::Model::shader() compiles, from the constructor parameters alone
    (pas_shader_source): needs neither a model nor a GPU -- the reference compiles it in its
    constructor, before Init (atmosphere/model.cc:769-776)."""
    lib = load_library()
    p, keep = _make_params(spec.wavelengths, spec.solar_irradiance, spec.sun_angular_radius, spec.bottom_radius,
                           spec.top_radius, spec.rayleigh_density, spec.rayleigh_scattering, spec.mie_density,
                           spec.mie_scattering, spec.mie_extinction, spec.mie_phase_function_g,
                           spec.absorption_density, spec.absorption_extinction, spec.ground_albedo,
                           spec.max_sun_zenith_angle, spec.length_unit_in_meters,
                           spec.num_precomputed_wavelengths, spec.combine_scattering_textures,
                           spec.half_precision, sizes)
    n = ctypes.c_size_t()
    d = glsl_directory.encode()
    _check(lib.pas_shader_source(ctypes.byref(p), d, None, ctypes.byref(n)))
    buf = ctypes.create_string_buffer(n.value)
    _check(lib.pas_shader_source(ctypes.byref(p), d, buf, ctypes.byref(n)))
    return buf.value.decode("utf-8")


class Model:
    """Same 19 constructor arguments as atmosphere::Model (atmosphere/model.h:182-281), plus the
    run-time extensions of the C ABI (`sizes`, `device`)."""

    kLambdaR, kLambdaG, kLambdaB = 680.0, 550.0, 440.0

    def __init__(self, wavelengths, solar_irradiance, sun_angular_radius, bottom_radius, top_radius,
                 rayleigh_density, rayleigh_scattering, mie_density, mie_scattering, mie_extinction,
                 mie_phase_function_g, absorption_density, absorption_extinction, ground_albedo,
                 max_sun_zenith_angle, length_unit_in_meters, num_precomputed_wavelengths,
                 combine_scattering_textures, half_precision, *, sizes: Optional[Dict[str, int]] = None,
                 device: Optional[int] = None):
        self._lib = load_library()
        self._h = ctypes.c_void_p()
        p, keep = _make_params(wavelengths, solar_irradiance, sun_angular_radius, bottom_radius, top_radius,
                               rayleigh_density, rayleigh_scattering, mie_density, mie_scattering,
                               mie_extinction, mie_phase_function_g, absorption_density, absorption_extinction,
                               ground_albedo, max_sun_zenith_angle, length_unit_in_meters,
                               num_precomputed_wavelengths, combine_scattering_textures, half_precision,
                               sizes, device)
        _check(self._lib.pas_model_create(ctypes.byref(p), ctypes.byref(self._h)))
        self.half_precision = bool(half_precision)
        self.combine_scattering_textures = bool(combine_scattering_textures)
        self.device = device

    @classmethod
    def from_spec(cls, spec: AtmosphereSpec, **kw) -> "Model":
        return cls(spec.wavelengths, spec.solar_irradiance, spec.sun_angular_radius, spec.bottom_radius,
                   spec.top_radius, spec.rayleigh_density, spec.rayleigh_scattering, spec.mie_density,
                   spec.mie_scattering, spec.mie_extinction, spec.mie_phase_function_g,
                   spec.absorption_density, spec.absorption_extinction, spec.ground_albedo,
                   spec.max_sun_zenith_angle, spec.length_unit_in_meters,
                   spec.num_precomputed_wavelengths, spec.combine_scattering_textures,
                   spec.half_precision, **kw)

    # -- lifetime ------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self._lib.pas_model_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- the reference API -----------------------------------------------------------------------
    def Init(self, num_scattering_orders: int = 4) -> None:
        """atmosphere::Model::Init (atmosphere/model.cc:866-975)."""
        _check(self._lib.pas_model_init(self._h, int(num_scattering_orders)))

    def InitAsync(self, num_scattering_orders: int = 4) -> None:
        """Enqueues Init on the model's own CUDA streams and returns (pas_model_init_async); several
        models can be in flight at once. ``Wait`` -- or any call that reads results -- blocks."""
        _check(self._lib.pas_model_init_async(self._h, int(num_scattering_orders)))

    def Wait(self) -> None:
        _check(self._lib.pas_model_wait(self._h))

    def set_host_outputs(self, transmittance: Optional[np.ndarray] = None, scattering: Optional[np.ndarray] = None,
                         single_mie_scattering: Optional[np.ndarray] = None,
                         irradiance: Optional[np.ndarray] = None) -> None:
        """Registers host arrays (native texel format: what ``texture(which, as_float32=False)``
        returns; pinned memory for overlap) that every later ``Init`` fills while it runs
        (pas_model_set_host_outputs). The arrays must stay alive until they are unregistered."""
        ptrs = []
        for which, arr in ((TEXTURE_TRANSMITTANCE, transmittance), (TEXTURE_SCATTERING, scattering),
                           (TEXTURE_SINGLE_MIE, single_mie_scattering), (TEXTURE_IRRADIANCE, irradiance)):
            if arr is None:
                ptrs.append(None)
                continue
            info = self.texture_info(which)
            need = info.width * info.height * info.depth * 4 * info.bytes_per_channel
            if not (arr.flags["C_CONTIGUOUS"] and arr.nbytes == need):
                raise ValueError(f"host output {which}: need a contiguous array of {need} bytes")
            ptrs.append(ctypes.c_void_p(arr.ctypes.data))
        self._host_outputs = (transmittance, scattering, single_mie_scattering, irradiance)  # keep alive
        _check(self._lib.pas_model_set_host_outputs(self._h, *ptrs))

    def set_host_output_mode(self, own_layers_only: bool) -> None:
        """Multi-GPU worlds: copy out only the layers this rank computed (the registered arrays are then
        host tables SHARED by the ranks, world.shared_host_tables) -- pas_model_set_host_output_mode."""
        _check(self._lib.pas_model_set_host_output_mode(self._h, 1 if own_layers_only else 0))

    def GetShaderSource(self, glsl_directory: str) -> str:
        """The source atmosphere::Model::shader() compiles (atmosphere/model.cc:691-744, 769-772)."""
        size = ctypes.c_size_t(0)
        _check(self._lib.pas_model_shader_source(self._h, glsl_directory.encode(), None, ctypes.byref(size)))
        buf = ctypes.create_string_buffer(size.value)
        _check(self._lib.pas_model_shader_source(self._h, glsl_directory.encode(), buf, ctypes.byref(size)))
        return buf.value.decode()

    # -- render-time API (atmosphere/reference/model.h:62-77; GLSL API of atmosphere/model.cc:221-281) --
    @staticmethod
    def _vec3(v, n=None) -> np.ndarray:
        a = np.ascontiguousarray(np.asarray(v, dtype=np.float64))
        if a.ndim == 1:
            a = np.ascontiguousarray(np.broadcast_to(a, ((n or 1), 3)))
        assert a.shape[-1] == 3
        return a

    def GetSolarRadiance(self, use_luminance: bool = False) -> np.ndarray:
        out = (ctypes.c_double * 3)()
        _check(self._lib.pas_model_get_solar_radiance(self._h, int(use_luminance), out))
        return np.array(list(out))

    def _sky(self, fn, camera, target, shadow_length, sun_direction, use_luminance):
        cam = self._vec3(camera)
        n = cam.shape[0]
        tgt, sun = self._vec3(target, n), self._vec3(sun_direction, n)
        sl = None
        if shadow_length is not None:
            sl = np.ascontiguousarray(np.broadcast_to(np.asarray(shadow_length, dtype=np.float64), (n,)))
        rad, tr = np.empty((n, 3), np.float32), np.empty((n, 3), np.float32)
        _check(fn(self._h, int(use_luminance), n, cam.ctypes.data, tgt.ctypes.data,
                  sl.ctypes.data if sl is not None else None, sun.ctypes.data, rad.ctypes.data, tr.ctypes.data))
        return rad, tr

    def GetSkyRadiance(self, camera, view_ray, shadow_length, sun_direction, use_luminance: bool = False):
        """Batched GetSkyRadiance (functions.glsl:1705-1769): arrays [n, 3]; returns (radiance,
        transmittance), float32 [n, 3]."""
        return self._sky(self._lib.pas_model_get_sky_radiance, camera, view_ray, shadow_length,
                         sun_direction, use_luminance)

    def GetSkyRadianceToPoint(self, camera, point, shadow_length, sun_direction, use_luminance: bool = False):
        """Batched GetSkyRadianceToPoint (functions.glsl:1787-1863)."""
        return self._sky(self._lib.pas_model_get_sky_radiance_to_point, camera, point, shadow_length,
                         sun_direction, use_luminance)

    def sky_radiance_device(self, n: int, camera_ptr: int, view_ray_ptr: int, shadow_length_ptr: int,
                            sun_direction_ptr: int, radiance_ptr: int, transmittance_ptr: int = 0,
                            use_luminance: bool = False) -> None:
        """pas_model_get_sky_radiance on DEVICE arrays (raw pointers: [n][3] float64 inputs, [n] float64
        shadow lengths or 0, [n][3] float32 outputs): nothing crosses PCIe."""
        p = lambda v: ctypes.c_void_p(int(v) or None)
        _check(self._lib.pas_model_get_sky_radiance(self._h, int(use_luminance), int(n), p(camera_ptr),
                                                   p(view_ray_ptr), p(shadow_length_ptr), p(sun_direction_ptr),
                                                   p(radiance_ptr), p(transmittance_ptr)))

    def GetSunAndSkyIrradiance(self, point, normal, sun_direction, use_luminance: bool = False):
        """Batched GetSunAndSkyIrradiance (functions.glsl:1878-1896): returns (sun, sky)."""
        p = self._vec3(point)
        n = p.shape[0]
        nr, sun = self._vec3(normal, n), self._vec3(sun_direction, n)
        e0, e1 = np.empty((n, 3), np.float32), np.empty((n, 3), np.float32)
        _check(self._lib.pas_model_get_sun_and_sky_irradiance(self._h, int(use_luminance), n, p.ctypes.data,
                                                             nr.ctypes.data, sun.ctypes.data, e0.ctypes.data,
                                                             e1.ctypes.data))
        return e0, e1

    def render_context(self, use_luminance: bool = False) -> bytes:
        """The pas::RenderContext of csrc/kernel_render.cuh for this model (pas_model_render_context): what
        a CUDA renderer passes to its own kernels to call the lookups on the device."""
        n = ctypes.c_size_t()
        _check(self._lib.pas_model_render_context(self._h, int(use_luminance), None, ctypes.byref(n)))
        buf = ctypes.create_string_buffer(n.value)
        _check(self._lib.pas_model_render_context(self._h, int(use_luminance), buf, ctypes.byref(n)))
        return buf.raw[:n.value]

    def last_render_ms(self) -> float:
        ms = ctypes.c_float(0)
        _check(self._lib.pas_model_last_render_ms(self._h, ctypes.byref(ms)))
        return ms.value

    # -- tables ----------------------------------------------------------------------------------
    def texture_info(self, which: int) -> _TextureInfo:
        info = _TextureInfo()
        _check(self._lib.pas_model_texture_info(self._h, which, ctypes.byref(info)))
        return info

    def texture(self, which: int, as_float32: bool = True, out: Optional[np.ndarray] = None) -> np.ndarray:
        """RGBA texels, shape (depth, height, width, 4) (2-D tables: (height, width, 4))."""
        info = self.texture_info(which)
        if not info.present:
            raise PasError(5, "this model has no such table")
        dtype = np.float32 if (as_float32 or info.bytes_per_channel == 4) else np.float16
        shape = ((info.depth,) if info.depth > 1 else ()) + (info.height, info.width, 4)
        if out is None:
            out = np.empty(shape, dtype=dtype)
        assert out.dtype == dtype and out.flags["C_CONTIGUOUS"] and out.size == int(np.prod(shape))
        _check(self._lib.pas_model_read_texture(self._h, which, int(as_float32), out.ctypes.data, out.nbytes))
        return out.reshape(shape)

    def device_ptr(self, which: int) -> int:
        p = ctypes.c_void_p()
        _check(self._lib.pas_model_texture_device_ptr(self._h, which, ctypes.byref(p)))
        return p.value

    @property
    def transmittance(self):
        return self.texture(TEXTURE_TRANSMITTANCE)

    @property
    def scattering(self):
        return self.texture(TEXTURE_SCATTERING)

    @property
    def irradiance(self):
        return self.texture(TEXTURE_IRRADIANCE)

    @property
    def single_mie_scattering(self):
        return self.texture(TEXTURE_SINGLE_MIE)

    def save_webgl(self, directory: str, glsl_directory: str, vertex_shader: Optional[str] = None,
                   fragment_shader: Optional[str] = None) -> None:
        """The WebGL hand-off of atmosphere/demo/webgl/precompute.cc:81-106: the .dat files,
        atmosphere_shader.txt and (the caller's own) vertex_shader.txt / fragment_shader.txt."""
        enc = lambda t: None if t is None else t.encode()
        _check(self._lib.pas_model_save_webgl(self._h, directory.encode(), glsl_directory.encode(),
                                              enc(vertex_shader), enc(fragment_shader)))

    def save_dat(self, directory: str) -> None:
        _check(self._lib.pas_model_save_dat(self._h, directory.encode()))

    # -- introspection ---------------------------------------------------------------------------
    def channels(self) -> List[float]:
        n = ctypes.c_int()
        _check(self._lib.pas_model_channels(self._h, ctypes.byref(n), None))
        lam = (ctypes.c_double * n.value)()
        _check(self._lib.pas_model_channels(self._h, ctypes.byref(n), lam))
        return list(lam)

    def luminance_matrix(self) -> np.ndarray:
        c = len(self.channels())
        out = np.empty((3, c), dtype=np.float32)
        _check(self._lib.pas_model_luminance_matrix(self._h, out.ctypes.data_as(ctypes.POINTER(ctypes.c_float))))
        return out

    def luminance_factors(self):
        out = (ctypes.c_double * 6)()
        _check(self._lib.pas_model_luminance_factors(self._h, out))
        return list(out[:3]), list(out[3:])

    def set_capture(self, enabled: bool) -> None:
        _check(self._lib.pas_model_set_capture(self._h, int(enabled)))

    def _table_shape(self, name: str):
        t, s, e = (self.texture_info(i) for i in (TEXTURE_TRANSMITTANCE, TEXTURE_SCATTERING, TEXTURE_IRRADIANCE))
        if name.startswith("transmittance"):
            return (t.height, t.width)
        if name.startswith("delta_irradiance"):
            return (e.height, e.width)
        return (s.depth, s.height, s.width)

    def intermediate(self, name: str) -> np.ndarray:
        """Planar per-channel copy [C, ...] of an intermediate (needs set_capture(True) before Init,
        or one of the live-buffer names after run_phase)."""
        n = ctypes.c_size_t(0)
        _check(self._lib.pas_model_read_intermediate(self._h, name.encode(), None, ctypes.byref(n)))
        out = np.empty(n.value, dtype=np.float32)
        _check(self._lib.pas_model_read_intermediate(self._h, name.encode(), out.ctypes.data, ctypes.byref(n)))
        shape = self._table_shape(name)
        return out.reshape((-1,) + shape)

    def write_intermediate(self, name: str, array: np.ndarray) -> None:
        a = np.ascontiguousarray(array, dtype=np.float32)
        _check(self._lib.pas_model_write_intermediate(self._h, name.encode(), a.ctypes.data, a.size))

    def run_phase(self, phase: str, order: int = 0) -> None:
        _check(self._lib.pas_model_run_phase(self._h, PHASES[phase], order))

    def last_timings(self) -> Dict[str, float]:
        n = ctypes.c_int(0)
        _check(self._lib.pas_model_last_timings(self._h, ctypes.byref(n), None, None))
        names = (ctypes.c_char_p * n.value)()
        ms = (ctypes.c_float * n.value)()
        _check(self._lib.pas_model_last_timings(self._h, ctypes.byref(n), names, ms))
        out: Dict[str, float] = {}
        for k, v in zip(names, ms):
            key = k.decode()
            out[key] = out.get(key, 0.0) + float(v)
        return out

    def last_launch_count(self) -> int:
        n = ctypes.c_int()
        _check(self._lib.pas_model_last_launch_count(self._h, ctypes.byref(n)))
        return n.value

    def attach_world(self, rank: int, world_size: int, unique_id: Optional[bytes]) -> None:
        _check(self._lib.pas_model_attach_world(self._h, rank, world_size, unique_id))

    def ipc_export(self, rank: int, world_size: int) -> bytes:
        """CUDA IPC handles of this rank's exchange buffers (pas_model_ipc_export), to be
        all-gathered by the host and handed to ``attach_peers`` on every rank."""
        n = ctypes.c_size_t(IPC_EXPORT_BYTES)
        buf = ctypes.create_string_buffer(IPC_EXPORT_BYTES)
        _check(self._lib.pas_model_ipc_export(self._h, rank, world_size, buf, ctypes.byref(n)))
        return buf.raw[:n.value]

    def exchange_bytes(self, world_size: int) -> int:
        """Size of the symmetric arena this model needs in a world of ``world_size`` ranks."""
        n = ctypes.c_size_t()
        _check(self._lib.pas_model_exchange_bytes(self._h, int(world_size), ctypes.byref(n)))
        return n.value

    def attach_symmetric(self, rank: int, world_size: int, arena_ptrs: Sequence[int], multicast_ptr: int,
                         arena_bytes: int) -> None:
        """pas_model_attach_symmetric: ``arena_ptrs[r]`` = rank r's arena mapped in this process,
        ``multicast_ptr`` = the NVLS multicast address of the arenas (0 = none)."""
        bases = (ctypes.c_void_p * world_size)(*[int(p) for p in arena_ptrs])
        _check(self._lib.pas_model_attach_symmetric(self._h, int(rank), int(world_size), bases,
                                                    ctypes.c_void_p(int(multicast_ptr) or None), int(arena_bytes)))

    def attach_peers(self, exports: bytes, bytes_per_rank: int = 0) -> None:
        """Maps the other ranks' buffers: ``exports`` = the ranks' ``ipc_export`` blobs in rank order."""
        _check(self._lib.pas_model_attach_peers(self._h, exports, bytes_per_rank or IPC_EXPORT_BYTES))

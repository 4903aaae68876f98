"""TEST INFRASTRUCTURE: ctypes binding of oracle/_ref/libpas_ref.so, the UNMODIFIED reference CPU
model (see oracle/ref_driver.cc). Only tests/, oracle/ scripts, __graft_entry__.smoke() and
bench.py's CPU arms may import this module; the product package never does."""
from __future__ import annotations

import ctypes
import os
from typing import Dict

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libpas_ref.so")

TABLES = {"transmittance": 0, "delta_irradiance": 1, "irradiance": 2, "delta_rayleigh": 3,
          "delta_mie": 4, "delta_density": 5, "delta_multiple": 6, "scattering": 7}
PHASES = {"transmittance": 0, "direct_irradiance": 1, "single_scattering": 2,
          "scattering_density": 3, "indirect_irradiance": 4, "multiple_scattering": 5}
MAX_LANES = 47


def available() -> bool:
    return os.path.exists(LIB_PATH)


def _lib():
    lib = ctypes.CDLL(LIB_PATH)
    lib.pasref_create.restype = ctypes.c_void_p
    lib.pasref_create.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    lib.pasref_destroy.argtypes = [ctypes.c_void_p]
    lib.pasref_phase.restype = ctypes.c_double
    lib.pasref_phase.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    lib.pasref_read.restype = ctypes.c_int
    lib.pasref_read.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    lib.pasref_sizes.argtypes = [ctypes.c_void_p]
    lib.pasref_uvwz_from_rmumusnu.argtypes = [ctypes.c_void_p] + [ctypes.c_double] * 4 + [ctypes.c_int, ctypes.c_void_p]
    lib.pasref_rmumusnu_from_frag_coord.argtypes = [ctypes.c_void_p] + [ctypes.c_double] * 3 + [ctypes.c_void_p]
    lib.pasref_write.restype = ctypes.c_int
    lib.pasref_write.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    lib.pasref_render_scene.restype = ctypes.c_double
    lib.pasref_render_scene.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                        ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    lib.pasref_sky_radiance.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                        ctypes.c_double, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    lib.pasref_sky_radiance_to_point.argtypes = lib.pasref_sky_radiance.argtypes
    lib.pasref_sun_and_sky_irradiance.argtypes = [ctypes.c_void_p, ctypes.c_int] + [ctypes.c_void_p] * 5
    return lib


def sizes() -> Dict[str, int]:
    out = (ctypes.c_int * 8)()
    _lib().pasref_sizes(out)
    keys = ["t_w", "t_h", "r", "mu", "mu_s", "nu", "e_w", "e_h"]
    return dict(zip(keys, list(out)))


class RefModel:
    """Reference CPU model with the C channels of `cp` (ChannelParams) packed into lanes 0..C-1."""

    def __init__(self, cp, nthreads: int | None = None):
        self.lib = _lib()
        self.n = cp.num_channels
        if self.n > MAX_LANES:
            raise ValueError("at most 47 lanes")
        self.nthreads = nthreads or os.cpu_count() or 1
        spectra = np.ascontiguousarray(np.stack([
            cp.solar_irradiance, cp.rayleigh_scattering, cp.mie_scattering, cp.mie_extinction,
            cp.absorption_extinction, cp.ground_albedo]).astype(np.float64))
        scalars = np.array([cp.sun_angular_radius, cp.bottom_radius, cp.top_radius,
                            cp.mie_phase_function_g, cp.mu_s_min], dtype=np.float64)
        profiles = np.ascontiguousarray(cp.profiles.astype(np.float64))
        self.h = self.lib.pasref_create(self.n, spectra.ctypes.data, scalars.ctypes.data,
                                        profiles.ctypes.data)
        if not self.h:
            raise RuntimeError("pasref_create failed")
        self.sz = sizes()

    def close(self):
        if self.h:
            self.lib.pasref_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def phase(self, name: str, order: int = 0, stride: int = 1) -> float:
        t = self.lib.pasref_phase(self.h, PHASES[name], order, stride, self.nthreads)
        if t < 0:
            raise RuntimeError(f"pasref_phase({name}) failed")
        return t

    def read(self, table: str) -> np.ndarray:
        """Planar copy [C, texels...] as float64, texel order x fastest."""
        s = self.sz
        shape = {0: (s["t_h"], s["t_w"]), 1: (s["e_h"], s["e_w"]), 2: (s["e_h"], s["e_w"])}.get(
            TABLES[table], (s["r"], s["mu"], s["nu"] * s["mu_s"]))
        out = np.empty((self.n,) + shape, dtype=np.float64)
        n = self.lib.pasref_read(self.h, TABLES[table], self.n, out.ctypes.data)
        assert n == int(np.prod(shape)), (n, shape)
        return out

    def write(self, table: str, array: np.ndarray) -> None:
        """Loads [C, texels...] float64 into lanes 0..C-1 of transmittance / irradiance / delta_mie
        (the single-Mie table of the render functions) / scattering."""
        a = np.ascontiguousarray(array, dtype=np.float64)
        assert a.shape[0] == self.n
        n = self.lib.pasref_write(self.h, TABLES[table], self.n, a.ctypes.data)
        assert n == a[0].size, (n, a.shape)

    def render_scene(self, view, ground_albedo, sphere_albedo, row_stride: int = 1) -> np.ndarray:
        """Spectral radiance [H, W, C] of the model_test.glsl scene through the reference's own
        functions. `view` is a precomputed_atmospheric_scattering_b200.scene.SceneView; albedos
        are per lane."""
        scene = np.array(list(view.camera) + list(view.earth_center) + list(view.sun_direction) +
                         list(view.sun_size) + list(view.sphere_center) + [view.sphere_radius] +
                         list(view.model_from_clip), dtype=np.float64)
        assert scene.size == 24
        alb = np.ascontiguousarray(np.concatenate([ground_albedo, sphere_albedo]), dtype=np.float64)
        assert alb.size == 2 * self.n
        out = np.zeros((view.height, view.width, self.n), dtype=np.float64)
        t = self.lib.pasref_render_scene(self.h, self.n, scene.ctypes.data, alb.ctypes.data, view.width,
                                         view.height, row_stride, self.nthreads, out.ctypes.data)
        if t < 0:
            raise RuntimeError("pasref_render_scene failed")
        self.last_render_seconds = t
        return out

    def _point(self, fn, a, b, shadow_length, sun):
        va, vb, vs = (np.array(v, dtype=np.float64) for v in (a, b, sun))
        o0, o1 = np.zeros(self.n), np.zeros(self.n)
        fn(self.h, self.n, va.ctypes.data, vb.ctypes.data, float(shadow_length), vs.ctypes.data,
           o0.ctypes.data, o1.ctypes.data)
        return o0, o1

    def sky_radiance(self, camera, view_ray, shadow_length, sun_direction):
        return self._point(self.lib.pasref_sky_radiance, camera, view_ray, shadow_length, sun_direction)

    def sky_radiance_to_point(self, camera, point, shadow_length, sun_direction):
        return self._point(self.lib.pasref_sky_radiance_to_point, camera, point, shadow_length, sun_direction)

    def sun_and_sky_irradiance(self, point, normal, sun_direction):
        vp, vn, vs = (np.array(v, dtype=np.float64) for v in (point, normal, sun_direction))
        o0, o1 = np.zeros(self.n), np.zeros(self.n)
        self.lib.pasref_sun_and_sky_irradiance(self.h, self.n, vp.ctypes.data, vn.ctypes.data,
                                               vs.ctypes.data, o0.ctypes.data, o1.ctypes.data)
        return o0, o1

    def uvwz(self, r, mu, mu_s, nu, hit):
        out = (ctypes.c_double * 4)()
        self.lib.pasref_uvwz_from_rmumusnu(self.h, r, mu, mu_s, nu, int(hit), out)
        return list(out)

    def rmumusnu(self, x, y, z):
        out = (ctypes.c_double * 5)()
        self.lib.pasref_rmumusnu_from_frag_coord(self.h, x, y, z, out)
        return list(out)

    def precompute(self, num_orders: int = 4, dump=None, log=None) -> Dict[str, float]:
        """The phase sequence of atmosphere/reference/model.cc:140-237. `dump(name, array)` is
        called with every intermediate as soon as it exists."""
        times: Dict[str, float] = {}

        def run(key, *a, **kw):
            times[key] = self.phase(*a, **kw)
            if log:
                log(f"{key}: {times[key]:.3f} s")

        def emit(name, table):
            if dump:
                dump(name, self.read(table))

        run("transmittance", "transmittance")
        emit("transmittance", "transmittance")
        run("direct_irradiance", "direct_irradiance")
        emit("delta_irradiance_1", "delta_irradiance")
        run("single_scattering", "single_scattering")
        emit("delta_rayleigh", "delta_rayleigh")
        emit("delta_mie", "delta_mie")
        for order in range(2, num_orders + 1):
            run(f"scattering_density_{order}", "scattering_density", order)
            emit(f"delta_density_{order}", "delta_density")
            run(f"indirect_irradiance_{order}", "indirect_irradiance", order)
            emit(f"delta_irradiance_{order}", "delta_irradiance")
            run(f"multiple_scattering_{order}", "multiple_scattering", order)
            emit(f"delta_multiple_{order}", "delta_multiple")
        emit("scattering", "scattering")
        emit("irradiance", "irradiance")
        return times

"""TEST INFRASTRUCTURE: ctypes binding of oracle/liboracle.so (oracle/pas_oracle.c), the fp64 CPU
restatement of the hot path, plus the host-side luminance/accumulation epilogues of the GL path
restated in numpy. Only tests/, oracle/ scripts, __graft_entry__.smoke() and bench.py's CPU arms
may import this module; the product package never does."""
from __future__ import annotations

import ctypes
import dataclasses
import math
import os
import subprocess
from typing import Callable, Dict, Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle.so")
MAX_CHANNELS = 48


@dataclasses.dataclass(frozen=True)
class Sizes:
    """Table sizes; the defaults are atmosphere/constants.h:47-61."""
    t_w: int = 256
    t_h: int = 64
    r: int = 32
    mu: int = 128
    mu_s: int = 32
    nu: int = 8
    e_w: int = 64
    e_h: int = 16

    @property
    def scattering_shape(self):
        return (self.r, self.mu, self.nu * self.mu_s)

    @property
    def rows3(self):
        return self.r * self.mu


class _CSizes(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int) for n in ("t_w", "t_h", "r", "mu", "mu_s", "nu", "e_w", "e_h")]


class _CAtm(ctypes.Structure):
    _fields_ = ([("nc", ctypes.c_int)] +
                [(n, ctypes.c_double * MAX_CHANNELS) for n in
                 ("solar_irradiance", "rayleigh_scattering", "mie_scattering", "mie_extinction",
                  "absorption_extinction", "ground_albedo")] +
                [(n, ctypes.c_double) for n in
                 ("sun_angular_radius", "bottom_radius", "top_radius", "mie_g", "mu_s_min")] +
                [("profiles", ctypes.c_double * 30), ("sz", _CSizes)])


class _CRenderTables(ctypes.Structure):
    _fields_ = ([(n, ctypes.c_void_p) for n in
                 ("transmittance", "scattering", "single_mie", "scattering_alpha", "irradiance")] +
                [("sky_k", ctypes.c_double * MAX_CHANNELS), ("sun_k", ctypes.c_double * MAX_CHANNELS),
                 ("gl_solar_radiance", ctypes.c_int)])


class _CScene(ctypes.Structure):
    _fields_ = [("camera", ctypes.c_double * 3), ("earth_center", ctypes.c_double * 3),
                ("sun_direction", ctypes.c_double * 3), ("sun_size", ctypes.c_double * 2),
                ("sphere_center", ctypes.c_double * 3), ("sphere_radius", ctypes.c_double),
                ("model_from_clip", ctypes.c_double * 9),
                ("ground_albedo", ctypes.c_double * MAX_CHANNELS),
                ("sphere_albedo", ctypes.c_double * MAX_CHANNELS),
                ("width", ctypes.c_int), ("height", ctypes.c_int)]


def build(force: bool = False) -> str:
    """Compiles liboracle.so with the recipe in oracle/Makefile (gcc only)."""
    srcs = [os.path.join(_HERE, f) for f in ("pas_oracle.c", "pas_oracle_render.c", "pas_oracle.h")]
    if force or not os.path.exists(LIB_PATH) or (
            os.path.getmtime(LIB_PATH) < max(os.path.getmtime(f) for f in srcs)):
        subprocess.check_call(["make", "-s", "-C", _HERE, LIB_PATH])
    return LIB_PATH


_lib_cache = None


def lib():
    global _lib_cache
    if _lib_cache is None:
        build()
        _lib_cache = ctypes.CDLL(LIB_PATH)
        l = _lib_cache
        for name in ("paso_distance_to_top", "paso_distance_to_bottom", "paso_profile_density",
                     "paso_optical_length_to_top", "paso_rayleigh_phase", "paso_mie_phase"):
            getattr(l, name).restype = ctypes.c_double
    return _lib_cache


def _p(arr: Optional[np.ndarray]):
    if arr is None:
        return ctypes.c_void_p(0)
    assert arr.dtype == np.float64 and arr.flags["C_CONTIGUOUS"]
    return ctypes.c_void_p(arr.ctypes.data)


class Oracle:
    """fp64 CPU oracle for the C channels of `cp` (atmospheres.ChannelParams)."""

    def __init__(self, cp, sizes: Sizes = Sizes()):
        self.cp, self.sz, self.nc = cp, sizes, cp.num_channels
        if self.nc > MAX_CHANNELS:
            raise ValueError("too many channels")
        a = _CAtm()
        a.nc = self.nc
        for name in ("solar_irradiance", "rayleigh_scattering", "mie_scattering", "mie_extinction",
                     "absorption_extinction", "ground_albedo"):
            arr = getattr(a, name)
            for i, v in enumerate(getattr(cp, name)):
                arr[i] = float(v)
        a.sun_angular_radius, a.bottom_radius, a.top_radius = (
            cp.sun_angular_radius, cp.bottom_radius, cp.top_radius)
        a.mie_g, a.mu_s_min = cp.mie_phase_function_g, cp.mu_s_min
        for i, v in enumerate(np.asarray(cp.profiles, dtype=np.float64).reshape(-1)):
            a.profiles[i] = float(v)
        for n in ("t_w", "t_h", "r", "mu", "mu_s", "nu", "e_w", "e_h"):
            setattr(a.sz, n, getattr(sizes, n))
        self.atm = a
        self.l = lib()

    # -- shapes ------------------------------------------------------------------------------
    def _t(self):
        return np.zeros((self.nc, self.sz.t_h, self.sz.t_w))

    def _e(self):
        return np.zeros((self.nc, self.sz.e_h, self.sz.e_w))

    def _s(self):
        return np.zeros((self.nc,) + self.sz.scattering_shape)

    def _rows(self, rows, full):
        return (0, full) if rows is None else rows

    # -- passes ------------------------------------------------------------------------------
    def transmittance(self, rows=None):
        T = self._t()
        b, e = self._rows(rows, self.sz.t_h)
        assert self.l.paso_transmittance(ctypes.byref(self.atm), _p(T), b, e) == 0
        return T

    def direct_irradiance(self, T, rows=None):
        dE = self._e()
        b, e = self._rows(rows, self.sz.e_h)
        assert self.l.paso_direct_irradiance(ctypes.byref(self.atm), _p(T), _p(dE), b, e) == 0
        return dE

    def single_scattering(self, T, rows=None):
        dR, dM = self._s(), self._s()
        b, e = self._rows(rows, self.sz.rows3)
        assert self.l.paso_single_scattering(ctypes.byref(self.atm), _p(T), _p(dR), _p(dM), b, e) == 0
        return dR, dM

    def scattering_density(self, T, dR, dM, dS, dE, order, rows=None):
        dJ = self._s()
        b, e = self._rows(rows, self.sz.rows3)
        assert self.l.paso_scattering_density(ctypes.byref(self.atm), _p(T), _p(dR), _p(dM),
                                              _p(dS), _p(dE), order, _p(dJ), b, e) == 0
        return dJ

    def indirect_irradiance(self, dR, dM, dS, order, rows=None):
        """`order` is the order of the radiance being integrated (the reference calls this with
        scattering_order - 1, reference/model.cc:210)."""
        dE = self._e()
        b, e = self._rows(rows, self.sz.e_h)
        assert self.l.paso_indirect_irradiance(ctypes.byref(self.atm), _p(dR), _p(dM), _p(dS),
                                               order, _p(dE), b, e) == 0
        return dE

    def multiple_scattering(self, T, dJ, rows=None):
        dS = self._s()
        nu = np.zeros(self.sz.scattering_shape)
        b, e = self._rows(rows, self.sz.rows3)
        assert self.l.paso_multiple_scattering(ctypes.byref(self.atm), _p(T), _p(dJ), _p(dS),
                                               _p(nu), b, e) == 0
        return dS, nu

    def texel_params(self):
        """(r, mu, mu_s, nu, hit) of every scattering texel, each shaped like the 3-D table."""
        out = np.zeros(self.sz.scattering_shape + (5,))
        buf = (ctypes.c_double * 5)()
        f = self.l.paso_rmumusnu_from_frag_coord
        f.argtypes = [ctypes.c_void_p, ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_void_p]
        R, MU, W = self.sz.scattering_shape
        for k in range(R):
            for j in range(MU):
                for i in range(W):
                    f(ctypes.byref(self.atm), i + 0.5, j + 0.5, k + 0.5, buf)
                    out[k, j, i] = buf[:]
        return out

    # -- the whole job -----------------------------------------------------------------------
    def precompute(self, num_orders: int = 4, dump: Optional[Callable] = None,
                   log: Optional[Callable] = None) -> Dict[str, np.ndarray]:
        """Phase sequence of atmosphere/reference/model.cc:140-237; returns every intermediate
        with the names oracle/ref.py uses."""
        import time
        out: Dict[str, np.ndarray] = {}

        def put(name, arr):
            out[name] = arr
            if dump:
                dump(name, arr)

        def timed(name, f, *a, **kw):
            t0 = time.time()
            r = f(*a, **kw)
            if log:
                log(f"{name}: {time.time() - t0:.3f} s")
            return r

        T = timed("transmittance", self.transmittance)
        put("transmittance", T)
        dE = timed("direct_irradiance", self.direct_irradiance, T)
        put("delta_irradiance_1", dE)
        dR, dM = timed("single_scattering", self.single_scattering, T)
        put("delta_rayleigh", dR)
        put("delta_mie", dM)
        S = dR.copy()
        E = np.zeros_like(dE)
        dS = np.zeros_like(dR)
        for order in range(2, num_orders + 1):
            dJ = timed(f"scattering_density_{order}", self.scattering_density, T, dR, dM, dS, dE, order)
            put(f"delta_density_{order}", dJ)
            dE = timed(f"indirect_irradiance_{order}", self.indirect_irradiance, dR, dM, dS, order - 1)
            put(f"delta_irradiance_{order}", dE)
            E = E + dE
            dS, nu = timed(f"multiple_scattering_{order}", self.multiple_scattering, T, dJ)
            put(f"delta_multiple_{order}", dS)
            S = S + dS / rayleigh_phase(nu)[None]
            out["nu"] = nu
        put("scattering", S)
        put("irradiance", E)
        return out

    # -- point functions (analytic known-answer tests, reference/functions_test.cc) ----------
    def _call(self, name, *args, nout=0, ret=None):
        """Calls paso_<name>(atm, *args[, out]) with doubles / ints / table pointers; returns the
        nout-vector written by the callee, or its scalar result."""
        f = getattr(self.l, "paso_" + name)
        cargs = [ctypes.byref(self.atm)]
        for a in args:
            if a is None or isinstance(a, np.ndarray):
                cargs.append(_p(a))
            elif isinstance(a, (bool, int, np.integer)):
                cargs.append(ctypes.c_int(int(a)))
            else:
                cargs.append(ctypes.c_double(float(a)))
        if nout:
            out = np.zeros(nout)
            cargs.append(_p(out))
            f.restype = None
            f(*cargs)
            return out
        f.restype = ret or ctypes.c_double
        return f(*cargs)

    def distance_to_top(self, r, mu):
        return self._call("distance_to_top", float(r), float(mu))

    def distance_to_bottom(self, r, mu):
        return self._call("distance_to_bottom", float(r), float(mu))

    def ray_intersects_ground(self, r, mu):
        return bool(self._call("ray_intersects_ground", float(r), float(mu), ret=ctypes.c_int))

    def profile_density(self, profile, altitude):
        return self._call("profile_density", int(profile), float(altitude))

    def optical_length_to_top(self, profile, r, mu):
        return self._call("optical_length_to_top", int(profile), float(r), float(mu))

    def compute_transmittance_to_top(self, r, mu):
        return self._call("compute_transmittance_to_top", float(r), float(mu), nout=self.nc)

    def transmittance_uv_from_rmu(self, r, mu):
        return self._call("transmittance_uv_from_rmu", float(r), float(mu), nout=2)

    def rmu_from_transmittance_uv(self, u, v):
        return self._call("rmu_from_transmittance_uv", float(u), float(v), nout=2)

    def scattering_uvwz_from_rmumusnu(self, r, mu, mu_s, nu, hit):
        return self._call("scattering_uvwz_from_rmumusnu", float(r), float(mu), float(mu_s), float(nu),
                          bool(hit), nout=4)

    def rmumusnu_from_scattering_uvwz(self, uvwz):
        return self._call("rmumusnu_from_scattering_uvwz", np.ascontiguousarray(uvwz, dtype=np.float64),
                          nout=5)

    def rmumusnu_from_frag_coord(self, x, y, z):
        return self._call("rmumusnu_from_frag_coord", float(x), float(y), float(z), nout=5)

    def irradiance_uv_from_rmus(self, r, mu_s):
        return self._call("irradiance_uv_from_rmus", float(r), float(mu_s), nout=2)

    def rmus_from_irradiance_uv(self, u, v):
        return self._call("rmus_from_irradiance_uv", float(u), float(v), nout=2)

    def get_transmittance(self, T, r, mu, d, hit):
        return self._call("get_transmittance", T, float(r), float(mu), float(d), bool(hit), nout=self.nc)

    def get_transmittance_to_sun(self, T, r, mu_s):
        return self._call("get_transmittance_to_sun", T, float(r), float(mu_s), nout=self.nc)

    def get_scattering(self, tab, r, mu, mu_s, nu, hit):
        return self._call("get_scattering", tab, float(r), float(mu), float(mu_s), float(nu), bool(hit), nout=self.nc)

    def get_irradiance(self, E, r, mu_s):
        return self._call("get_irradiance", E, float(r), float(mu_s), nout=self.nc)

    def single_scattering_point(self, T, r, mu, mu_s, nu, hit):
        f = self.l.paso_single_scattering_point
        ray, mie = np.zeros(self.nc), np.zeros(self.nc)
        f.restype = None
        f(ctypes.byref(self.atm), _p(T), *(ctypes.c_double(float(v)) for v in (r, mu, mu_s, nu)),
          ctypes.c_int(int(bool(hit))), _p(ray), _p(mie))
        return ray, mie

    def scattering_density_point(self, T, dR, dM, dS, dE, r, mu, mu_s, nu, order):
        return self._call("scattering_density_point", T, dR, dM, dS, dE, float(r), float(mu), float(mu_s), float(nu), int(order),
                          nout=self.nc)

    def multiple_scattering_point(self, T, dJ, r, mu, mu_s, nu, hit):
        return self._call("multiple_scattering_point", T, dJ, float(r), float(mu), float(mu_s), float(nu), bool(hit), nout=self.nc)

    def indirect_irradiance_point(self, dR, dM, dS, r, mu_s, order):
        return self._call("indirect_irradiance_point", dR, dM, dS, float(r), float(mu_s), int(order), nout=self.nc)

    def direct_irradiance_point(self, T, r, mu_s):
        return self._call("direct_irradiance_point", T, float(r), float(mu_s), nout=self.nc)


class Renderer:
    """fp64 restatement of the render-time lookups and of the model_test.glsl scene
    (oracle/pas_oracle_render.c) over planar float64 tables [C, ...].

    `single_mie` None selects the combined-texture path: `scattering_alpha` [1, r, mu, w] then holds
    the red single Mie channel (functions.glsl:1625-1646). sky_k / sun_k are the luminance factors
    (1 for radiance output); gl_solar_radiance picks the GL model's solar radiance formula."""

    def __init__(self, oracle: "Oracle", transmittance, scattering, irradiance, single_mie=None,
                 scattering_alpha=None, sky_k=None, sun_k=None, gl_solar_radiance=True):
        self.o, self.nc = oracle, oracle.nc
        f64 = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float64)
        self._keep = [f64(transmittance), f64(scattering), f64(single_mie), f64(scattering_alpha), f64(irradiance)]
        assert self._keep[2] is not None or self._keep[3] is not None
        t = _CRenderTables()
        for name, arr in zip(("transmittance", "scattering", "single_mie", "scattering_alpha", "irradiance"),
                             self._keep):
            setattr(t, name, None if arr is None else arr.ctypes.data)
        for c in range(self.nc):
            t.sky_k[c] = 1.0 if sky_k is None else float(sky_k[c])
            t.sun_k[c] = 1.0 if sun_k is None else float(sun_k[c])
        t.gl_solar_radiance = int(bool(gl_solar_radiance))
        self.tab = t
        self.l = oracle.l

    def _scene(self, view, ground_albedo=None, sphere_albedo=None) -> _CScene:
        s = _CScene()
        for name in ("camera", "earth_center", "sun_direction", "sun_size", "sphere_center", "model_from_clip"):
            getattr(s, name)[:] = list(getattr(view, name))
        s.sphere_radius = view.sphere_radius
        ga = view.ground_albedo if ground_albedo is None else ground_albedo
        sa = view.sphere_albedo if sphere_albedo is None else sphere_albedo
        assert len(ga) == self.nc and len(sa) == self.nc
        for c in range(self.nc):
            s.ground_albedo[c], s.sphere_albedo[c] = float(ga[c]), float(sa[c])
        s.width, s.height = view.width, view.height
        return s

    def render_scene(self, view, ground_albedo=None, sphere_albedo=None, threads: Optional[int] = None) -> np.ndarray:
        """[H, W, C] float64 radiance (or luminance) before tone mapping."""
        from concurrent.futures import ThreadPoolExecutor
        s = self._scene(view, ground_albedo, sphere_albedo)
        out = np.zeros((view.height, view.width, self.nc))
        n = threads or os.cpu_count() or 1
        bounds = np.linspace(0, view.height, min(n, view.height) + 1).astype(int)

        def job(i):
            return self.l.paso_render_scene(ctypes.byref(self.o.atm), ctypes.byref(self.tab), ctypes.byref(s),
                                            _p(out), int(bounds[i]), int(bounds[i + 1]))
        with ThreadPoolExecutor(len(bounds) - 1) as ex:
            assert all(r == 0 for r in ex.map(job, range(len(bounds) - 1)))
        return out

    def _two(self, name, a, b, shadow_length, sun):
        va, vb, vs = (np.array(v, dtype=np.float64) for v in (a, b, sun))
        o0, o1 = np.zeros(self.nc), np.zeros(self.nc)
        getattr(self.l, name)(ctypes.byref(self.o.atm), ctypes.byref(self.tab), _p(va), _p(vb),
                              ctypes.c_double(float(shadow_length)), _p(vs), _p(o0), _p(o1))
        return o0, o1

    def sky_radiance(self, camera, view_ray, shadow_length, sun_direction):
        return self._two("paso_sky_radiance", camera, view_ray, shadow_length, sun_direction)

    def sky_radiance_to_point(self, camera, point, shadow_length, sun_direction):
        return self._two("paso_sky_radiance_to_point", camera, point, shadow_length, sun_direction)

    def sun_and_sky_irradiance(self, point, normal, sun_direction):
        vp, vn, vs = (np.array(v, dtype=np.float64) for v in (point, normal, sun_direction))
        o0, o1 = np.zeros(self.nc), np.zeros(self.nc)
        self.l.paso_sun_and_sky_irradiance(ctypes.byref(self.o.atm), ctypes.byref(self.tab), _p(vp), _p(vn),
                                           _p(vs), _p(o0), _p(o1))
        return o0, o1

    def solar_radiance(self):
        out = np.zeros(self.nc)
        self.l.paso_solar_radiance.restype = None
        self.l.paso_solar_radiance(ctypes.byref(self.o.atm), ctypes.byref(self.tab), _p(out))
        return out


def rayleigh_phase(nu):
    """atmosphere/functions.glsl:739-742."""
    return 3.0 / (16.0 * math.pi) * (1.0 + nu * nu)


# ---- host-side pieces of the GL path (atmosphere/model.cc), restated for the final tables ------

def luminance_from_radiance(lambdas: Sequence[float], cie_table: np.ndarray,
                            xyz_to_srgb: np.ndarray) -> np.ndarray:
    """L[3][C] with L[c][j] = float32((XYZ_TO_SRGB . cie(lambda_j))_c * dlambda), the matrices of
    atmosphere/model.cc:925-943 laid side by side for all batches (dlambda = 470 / C). The CIE
    table / matrix are passed in by the caller (tests read them from the product's host library so
    that no reference data is restated here)."""
    C = len(lambdas)
    dl = (830.0 - 360.0) / C
    L = np.zeros((3, C))
    for j, lam in enumerate(lambdas):
        xyz = np.array([cie_value(cie_table, lam, col) for col in (1, 2, 3)])
        L[:, j] = np.float32((xyz_to_srgb.reshape(3, 3) @ xyz) * dl)
    return L


def cie_value(cie_table: np.ndarray, wavelength: float, column: int) -> float:
    """atmosphere/model.cc:521-533: linear interpolation in the 5 nm CIE table, 0 outside."""
    if wavelength <= 360.0 or wavelength >= 830.0:
        return 0.0
    u = (wavelength - 360.0) / 5.0
    row = int(math.floor(u))
    u -= row
    t = cie_table.reshape(-1, 4)
    return t[row, column] * (1.0 - u) + t[row + 1, column] * u


def final_tables(inter: Dict[str, np.ndarray], L: np.ndarray, num_orders: int):
    """The accumulated products of the GL path from fp64 per-channel intermediates
    (atmosphere/model.cc:142-157, 176-208, 1082-1210): returns (S_rgb, S_alpha, M_rgb, E_rgb) with
    S_rgb = L.dR + sum_n L.dS_n / P_R(nu), S_alpha = (L.dM)_red, M_rgb = L.dM, E_rgb = sum L.dE_n
    (n >= 2; direct irradiance is not part of E, model.cc:139)."""
    mat = lambda X: np.tensordot(L, X, axes=(1, 0))
    S = mat(inter["delta_rayleigh"])
    M = mat(inter["delta_mie"])
    E = np.zeros((3,) + inter["delta_irradiance_1"].shape[1:])
    for n in range(2, num_orders + 1):
        S = S + mat(inter[f"delta_multiple_{n}"]) / rayleigh_phase(inter["nu"])[None]
        E = E + mat(inter[f"delta_irradiance_{n}"])
    return S, M[0], M, E

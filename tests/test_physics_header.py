"""csrc/pas_physics.cuh on the CPU: the header that holds every table mapping and geometric helper of
the kernels is compiled with g++ (tests/emu/physics_host.cc) and checked against the fp64 oracle --
the texel -> (r, mu, mu_s, nu) inverse maps, the forward maps in texel space, the boundary distances,
the density profiles, the index / weight / clamp rule of the software interpolation, and the fp32
inner-loop forms (rationalised distance to the top boundary, mu_s texel coordinate) within the
accuracy the kernels rely on. No GPU."""
import ctypes
import math
import os
import subprocess

import numpy as np
import pytest

import precomputed_atmospheric_scattering_b200 as pas

HERE = os.path.dirname(os.path.abspath(__file__))
SIZES = dict(t_w=64, t_h=16, r=8, mu=32, mu_s=8, nu=8, e_w=16, e_h=4)


class _Sizes(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int) for n in ("t_w", "t_h", "r_n", "mu_n", "mu_s_n", "nu_n", "e_w", "e_h")]


class _Geometry(ctypes.Structure):      # PasGeometry of csrc/pas_types.h
    _fields_ = [("sz", _Sizes), ("bottom", ctypes.c_double), ("top", ctypes.c_double), ("H", ctypes.c_double),
                ("mu_s_min", ctypes.c_double), ("mus_A", ctypes.c_double), ("sun_angular_radius", ctypes.c_double),
                ("mie_g", ctypes.c_double), ("profiles", ctypes.c_double * 30)]


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("emu") / "libphysics_host.so")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-ffp-contract=off", "-fPIC", "-shared", "-x", "c++",
                           os.path.join(HERE, "emu", "physics_host.cc"), "-o", out])
    lib = ctypes.CDLL(out)
    assert lib.emu_sizeof_geometry() == ctypes.sizeof(_Geometry)
    for name in ("emu_y_from_mu", "emu_x_from_mu_s", "emu_dist_top", "emu_dist_bottom", "emu_profile_density"):
        getattr(lib, name).restype = ctypes.c_double
    for name in ("emu_f_mu_s_texel_x", "emu_f_dist_top"):
        getattr(lib, name).restype = ctypes.c_float
    lib.emu_y_from_mu.argtypes = [ctypes.c_void_p, ctypes.c_double, ctypes.c_double, ctypes.c_int]
    lib.emu_x_from_mu_s.argtypes = [ctypes.c_void_p, ctypes.c_double]
    lib.emu_dist_top.argtypes = lib.emu_dist_bottom.argtypes = [ctypes.c_void_p, ctypes.c_double, ctypes.c_double]
    lib.emu_hits_ground.argtypes = [ctypes.c_void_p, ctypes.c_double, ctypes.c_double]
    lib.emu_profile_density.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_double]
    lib.emu_f_mu_s_texel_x.argtypes = [ctypes.c_void_p, ctypes.c_float]
    lib.emu_f_dist_top.argtypes = [ctypes.c_void_p, ctypes.c_double, ctypes.c_double]
    lib.emu_transmittance_xy.argtypes = [ctypes.c_void_p, ctypes.c_double, ctypes.c_double, ctypes.c_void_p]
    lib.emu_texel.argtypes = [ctypes.c_void_p] + [ctypes.c_int] * 4 + [ctypes.c_void_p]
    lib.emu_make_tap.argtypes = [ctypes.c_double, ctypes.c_int] + [ctypes.c_void_p] * 3
    lib.emu_make_tap_f.argtypes = [ctypes.c_float, ctypes.c_int] + [ctypes.c_void_p] * 3
    return lib


@pytest.fixture(scope="module", params=["small_planet", "earth"])
def setup(request, orc):
    spec = pas.small_planet() if request.param == "small_planet" else pas.earth(3, max_sun_zenith_deg=102.0)
    cp = pas.channel_params(spec, [680.0, 550.0, 440.0])
    o = orc.Oracle(cp, orc.Sizes(**SIZES))
    g = _Geometry()
    for a, b in (("t_w", "t_w"), ("t_h", "t_h"), ("r_n", "r"), ("mu_n", "mu"), ("mu_s_n", "mu_s"), ("nu_n", "nu"),
                 ("e_w", "e_w"), ("e_h", "e_h")):
        setattr(g.sz, a, SIZES[b])
    g.bottom, g.top = cp.bottom_radius, cp.top_radius
    g.H = math.sqrt(g.top * g.top - g.bottom * g.bottom)
    g.mu_s_min, g.sun_angular_radius, g.mie_g = cp.mu_s_min, cp.sun_angular_radius, cp.mie_phase_function_g
    d_min, d_max = g.top - g.bottom, g.H                         # "A" of functions.glsl:819-821
    g.mus_A = (o.distance_to_top(g.bottom, g.mu_s_min) - d_min) / (d_max - d_min)
    for i, v in enumerate(np.asarray(cp.profiles, dtype=np.float64).reshape(-1)):
        g.profiles[i] = float(v)
    return o, cp, g


def test_texel_inverse_maps_match_the_oracle(emu, setup):
    o, cp, g = setup
    out = (ctypes.c_double * 5)()
    for k in range(SIZES["r"]):
        for j in range(0, SIZES["mu"], 3):
            for i_mu_s in range(SIZES["mu_s"]):
                for i_nu in (0, 3, 7):
                    emu.emu_texel(ctypes.byref(g), k, j, i_mu_s, i_nu, out)
                    want = o.rmumusnu_from_frag_coord(i_nu * SIZES["mu_s"] + i_mu_s + 0.5, j + 0.5, k + 0.5)
                    assert np.allclose(out[:4], want[:4], rtol=1e-12, atol=1e-12), (k, j, i_mu_s, i_nu)
                    assert bool(out[4]) == bool(want[4])


def test_forward_maps_in_texel_space(emu, setup):
    o, cp, g = setup
    rng = np.random.default_rng(5)
    xy = (ctypes.c_double * 2)()
    for _ in range(300):
        r = cp.bottom_radius + rng.uniform(0.0, 1.0) * (cp.top_radius - cp.bottom_radius)
        mu, mu_s = rng.uniform(-1, 1), rng.uniform(cp.mu_s_min, 1)
        hit = bool(o.ray_intersects_ground(r, mu))
        u = o.scattering_uvwz_from_rmumusnu(r, mu, mu_s, 0.0, hit)          # (u_nu, u_mu_s, u_mu, u_r)
        assert emu.emu_y_from_mu(ctypes.byref(g), r, mu, int(hit)) == pytest.approx(u[2] * SIZES["mu"] - 0.5, abs=1e-9)
        assert emu.emu_x_from_mu_s(ctypes.byref(g), mu_s) == pytest.approx(u[1] * SIZES["mu_s"] - 0.5, abs=1e-9)
        emu.emu_transmittance_xy(ctypes.byref(g), r, mu, xy)
        uv = o.transmittance_uv_from_rmu(r, mu)
        assert xy[0] == pytest.approx(uv[0] * SIZES["t_w"] - 0.5, abs=1e-9)
        assert xy[1] == pytest.approx(uv[1] * SIZES["t_h"] - 0.5, abs=1e-9)
        assert emu.emu_dist_top(ctypes.byref(g), r, mu) == pytest.approx(o.distance_to_top(r, mu), rel=1e-13, abs=1e-12)
        assert bool(emu.emu_hits_ground(ctypes.byref(g), r, mu)) == hit
        if hit:
            assert emu.emu_dist_bottom(ctypes.byref(g), r, mu) == pytest.approx(o.distance_to_bottom(r, mu), rel=1e-13, abs=1e-12)
        h = r - cp.bottom_radius
        for p in range(3):
            assert emu.emu_profile_density(ctypes.byref(g), p, h) == pytest.approx(o.profile_density(p, h), rel=1e-14, abs=1e-300)


def test_fp32_inner_loop_forms(emu, setup):
    """The per-sample / per-direction forms run in fp32 on the device: texel coordinates must stay
    within ~1e-4 texel of the fp64 maps (a weight error of 1e-4 on neighbouring table values that
    differ by a few percent is far below the 1e-3 budget), also where the as-written expressions
    cancel (sun near the horizon, r near the top boundary)."""
    o, cp, g = setup
    rng = np.random.default_rng(6)
    for _ in range(500):
        mu_s = rng.uniform(cp.mu_s_min, 1.0)
        want = min(max(emu.emu_x_from_mu_s(ctypes.byref(g), mu_s), 0.0), SIZES["mu_s"] - 1.0)
        got = min(max(emu.emu_f_mu_s_texel_x(ctypes.byref(g), mu_s), 0.0), SIZES["mu_s"] - 1.0)
        assert abs(got - want) < 2e-4 * SIZES["mu_s"]
        r = cp.bottom_radius + rng.uniform(0.0, 1.0) ** 2 * (cp.top_radius - cp.bottom_radius)
        mu = rng.uniform(-1, 1)
        d = emu.emu_dist_top(ctypes.byref(g), r, mu)
        assert emu.emu_f_dist_top(ctypes.byref(g), r, mu) == pytest.approx(d, rel=2e-5, abs=2e-5 * (cp.top_radius - cp.bottom_radius))


def test_interpolation_taps_follow_the_reference_rule(emu):
    """dimensional_types/math/binary_function.h:103-118: i = floor(u n - 0.5), both indices clamped
    to [0, n - 1], weight = fractional part."""
    i0, i1, w = ctypes.c_int(), ctypes.c_int(), ctypes.c_float()
    for n in (2, 5, 32):
        for x in (-0.75, -0.5, 0.0, 0.25, 1.0, n - 1.5, n - 1.0, n - 0.6):
            emu.emu_make_tap(x, n, ctypes.byref(i0), ctypes.byref(i1), ctypes.byref(w))
            fl = math.floor(x)
            assert i0.value == min(max(fl, 0), n - 1) and i1.value == min(max(fl + 1, 0), n - 1)
            assert w.value == pytest.approx(x - fl, abs=1e-7)
            if 0.0 <= x <= n - 1:
                emu.emu_make_tap_f(x, n, ctypes.byref(i0), ctypes.byref(i1), ctypes.byref(w))
                # fp32 variant: same interpolated value (at x = n - 1 it returns (n - 2, n - 1, 1))
                v = lambda k: 10.0 + 3.0 * k
                got = v(i0.value) * (1 - w.value) + v(i1.value) * w.value
                assert got == pytest.approx(v(x), abs=1e-5)


def test_unclamped_texels_sit_exactly_on_a_nu_slab(emu, orc):
    """multiple_scattering_rows_kernel takes a 2-texel path for texels whose nu lies exactly on a slab
    of the source table: at the reference's sizes that is every texel whose nu
    was not clamped by functions.glsl:923-925 -- about 70 % -- because the slab value survives the
    round trip nu -> texel coordinate with a tap weight of exactly 0 or 1 in fp32."""
    sizes = dict(t_w=256, t_h=64, r=32, mu=128, mu_s=32, nu=8, e_w=64, e_h=16)
    spec = pas.earth(3, max_sun_zenith_deg=102.0)
    cp = pas.channel_params(spec, [680.0, 550.0, 440.0])
    o = orc.Oracle(cp, orc.Sizes(**sizes))
    g = _Geometry()
    for a, b in (("t_w", "t_w"), ("t_h", "t_h"), ("r_n", "r"), ("mu_n", "mu"), ("mu_s_n", "mu_s"), ("nu_n", "nu"),
                 ("e_w", "e_w"), ("e_h", "e_h")):
        setattr(g.sz, a, sizes[b])
    g.bottom, g.top = cp.bottom_radius, cp.top_radius
    g.H = math.sqrt(g.top * g.top - g.bottom * g.bottom)
    g.mu_s_min = cp.mu_s_min
    g.mus_A = (o.distance_to_top(g.bottom, g.mu_s_min) - (g.top - g.bottom)) / (g.H - (g.top - g.bottom))
    out = (ctypes.c_double * 5)()
    i0, i1, w = ctypes.c_int(), ctypes.c_int(), ctypes.c_float()
    on_slab = unclamped = total = 0
    for k in range(0, 32, 5):
        for j in range(0, 128, 7):
            for i_mu_s in range(32):
                for i_nu in range(8):
                    emu.emu_texel(ctypes.byref(g), k, j, i_mu_s, i_nu, out)
                    nu = out[3]
                    emu.emu_make_tap((nu + 1.0) * 0.5 * 7, 8, ctypes.byref(i0), ctypes.byref(i1), ctypes.byref(w))
                    slab = w.value == 0.0 or w.value == 1.0 or i0.value == i1.value
                    knot = i_nu / 7.0 * 2.0 - 1.0
                    total += 1
                    on_slab += slab
                    if nu == max(-1.0, min(1.0, knot)):          # not clamped by (mu, mu_s)
                        unclamped += 1
                        assert slab, (k, j, i_mu_s, i_nu, nu, w.value)
                        assert (i1.value if w.value == 1.0 else i0.value) == i_nu
    assert 0.6 < unclamped / total < 0.8 and on_slab >= unclamped

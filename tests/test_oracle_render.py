"""Pins the render-time restatement (oracle/pas_oracle_render.c) against the UNMODIFIED reference:
its GetSkyRadiance / GetSkyRadianceToPoint / GetSunAndSkyIrradiance (atmosphere/functions.glsl:1705-1896
compiled as C++ by atmosphere/reference/functions.cc) and its test scene
(atmosphere/reference/model_test.glsl included unmodified by oracle/ref_driver.cc), both evaluated on
the SAME tables, at the reference's table sizes. The tables are the reference's own transmittance
and single-scattering tables (4 s of CPU) -- the render functions do not care which order the
tables hold. CPU only; the reference-backed tests are skipped where oracle/_ref is absent, the
analytic ones always run.
"""
import math

import numpy as np
import pytest

from tests import scene

from oracle import ref

needs_ref = pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libpas_ref.so not built")
LAM = [680.0, 550.0, 440.0]


@pytest.fixture(scope="module")
def shared(pas, orc):
    """Reference model with T, single scattering and a non-trivial irradiance table + an oracle
    renderer over the same tables."""
    spec = pas.model_test_earth()
    cp = pas.channel_params(spec, LAM)
    model = ref.RefModel(cp)
    model.phase("transmittance")
    model.phase("direct_irradiance")
    model.phase("single_scattering")
    T, S, M = model.read("transmittance"), model.read("scattering"), model.read("delta_mie")
    E = model.read("delta_irradiance") * 0.05   # stands in for the sky irradiance
    model.write("irradiance", E)
    o = orc.Oracle(cp)
    r = orc.Renderer(o, T, S, E, single_mie=M, gl_solar_radiance=False)
    yield pas, spec, model, o, r, (T, S, M, E)
    model.close()


def rel(a, b, floor=1e-12):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor * max(np.abs(b).max(), 1e-300))))


@needs_ref
@pytest.mark.parametrize("zenith", [65.0, 88.0, 100.0])
def test_scene_matches_the_reference_renderer(shared, zenith):
    pas, spec, model, o, r, _ = shared
    view = scene.model_test_view(zenith, 90.0, False, width=96, height=54,
                                     sun_angular_radius=spec.sun_angular_radius)
    want = model.render_scene(view, view.ground_albedo, view.sphere_albedo)
    got = r.render_scene(view)
    assert np.isfinite(want).all() and want.max() > 0
    # same algorithm, same tables, fp64 on both sides
    assert rel(got, want, floor=1e-9) < 1e-9


@needs_ref
def test_point_lookups_match_the_reference(shared):
    pas, spec, model, o, r, _ = shared
    rng = np.random.default_rng(7)
    bottom, top = o.cp.bottom_radius, o.cp.top_radius
    worst = 0.0
    for trial in range(300):
        # cameras inside, at the top of, and outside the atmosphere
        radius = [bottom + rng.uniform(0.001, 59.0), bottom + 1e-3, top + rng.uniform(1.0, 3000.0)][trial % 3]
        d = rng.normal(size=3); d /= np.linalg.norm(d)
        camera = d * radius
        v = rng.normal(size=3); v /= np.linalg.norm(v)
        if trial % 3 == 2:
            v = -d + 0.02 * rng.normal(size=3); v /= np.linalg.norm(v)   # look at the planet
        sun = rng.normal(size=3); sun /= np.linalg.norm(sun)
        shadow = [0.0, rng.uniform(0.0, 20.0)][trial % 2]
        a, b = model.sky_radiance(camera, v, shadow, sun), r.sky_radiance(camera, v, shadow, sun)
        for x, y in zip(a, b):
            assert np.isnan(x).sum() == np.isnan(y).sum()
            if not np.isnan(x).any():
                worst = max(worst, rel(y, x, 1e-9))
        # a point on the view ray, inside the atmosphere
        dist = rng.uniform(0.1, 200.0)
        point = camera + v * dist
        if np.linalg.norm(point) < bottom or trial % 3 == 2:
            continue
        a = model.sky_radiance_to_point(camera, point, min(shadow, dist), sun)
        b = r.sky_radiance_to_point(camera, point, min(shadow, dist), sun)
        for x, y in zip(a, b):
            if not (np.isnan(x).any() or np.isnan(y).any()):
                worst = max(worst, rel(y, x, 1e-9))
        normal = rng.normal(size=3); normal /= np.linalg.norm(normal)
        p = d * (bottom + rng.uniform(0.0, 60.0))
        a, b = model.sun_and_sky_irradiance(p, normal, sun), r.sun_and_sky_irradiance(p, normal, sun)
        for x, y in zip(a, b):
            worst = max(worst, rel(y, x, 1e-9))
    assert worst < 1e-9, worst


def uniform_renderer(pas, orc, combined):
    """Tables that are constant in every texel: lookups then return the constants whatever the
    coordinates, which turns the render functions into closed forms."""
    spec = pas.model_test_earth()
    cp = pas.channel_params(spec, LAM)
    sz = orc.Sizes(t_w=8, t_h=4, r=2, mu=4, mu_s=2, nu=2, e_w=4, e_h=2)
    o = orc.Oracle(cp, sz)
    s_val, m_val = np.array([0.3, 0.5, 0.9]), np.array([0.02, 0.03, 0.05])
    T = np.full((3, sz.t_h, sz.t_w), 0.5)
    S = np.broadcast_to(s_val[:, None, None, None], (3,) + sz.scattering_shape).copy()
    M = np.broadcast_to(m_val[:, None, None, None], (3,) + sz.scattering_shape).copy()
    E = np.full((3, sz.e_h, sz.e_w), 0.2)
    if combined:
        r = orc.Renderer(o, T, S, E, single_mie=None, scattering_alpha=M[:1].copy())
    else:
        r = orc.Renderer(o, T, S, E, single_mie=M)
    return cp, o, r, s_val, m_val


@pytest.mark.parametrize("combined", [False, True])
def test_sky_radiance_closed_form_on_uniform_tables(pas, orc, combined):
    cp, o, r, s_val, m_val = uniform_renderer(pas, orc, combined)
    camera = np.array([0.0, 0.0, cp.bottom_radius + 1.0])
    view = np.array([0.0, math.sin(0.3), math.cos(0.3)])
    sun = np.array([math.sin(1.0), 0.0, math.cos(1.0)])
    nu = float(view @ sun)
    L, tr = r.sky_radiance(camera, view, 0.0, sun)
    mie = m_val
    if combined:
        # GetExtrapolatedSingleMieScattering (functions.glsl:1634-1646)
        mie = s_val * m_val[0] / s_val[0] * (cp.rayleigh_scattering[0] / cp.mie_scattering[0]) * (
            cp.mie_scattering / cp.rayleigh_scattering)
    want = s_val * o.l.paso_rayleigh_phase(ctypes_double(nu)) + mie * o.l.paso_mie_phase(
        ctypes_double(cp.mie_phase_function_g), ctypes_double(nu))
    assert np.allclose(L, want, rtol=1e-12)
    assert np.allclose(tr, 0.5)
    # a viewer in space looking away from the planet sees nothing (functions.glsl:1722-1726)
    L, tr = r.sky_radiance(np.array([0.0, 0.0, cp.top_radius + 100.0]), np.array([0.0, 0.0, 1.0]), 0.0, sun)
    assert np.all(L == 0.0) and np.all(tr == 1.0)
    # a ray into the ground has zero transmittance (functions.glsl:1733-1735)
    L, tr = r.sky_radiance(camera, np.array([0.0, 0.0, -1.0]), 0.0, sun)
    assert np.all(tr == 0.0)


def ctypes_double(v):
    import ctypes
    return ctypes.c_double(float(v))


def test_sun_and_sky_irradiance_closed_form(pas, orc):
    cp, o, r, _, _ = uniform_renderer(pas, orc, False)
    point = np.array([0.0, 0.0, cp.bottom_radius + 10.0])
    normal = np.array([0.0, 0.0, 1.0])
    sun = np.array([0.0, math.sin(0.5), math.cos(0.5)])
    e_sun, e_sky = r.sun_and_sky_irradiance(point, normal, sun)
    # horizontal surface: sky factor (1 + 1) / 2, sun fully above the horizon (functions.glsl:1878-1896)
    assert np.allclose(e_sky, 0.2)
    assert np.allclose(e_sun, cp.solar_irradiance * 0.5 * math.cos(0.5), rtol=1e-12)
    # a vertical surface sees half the sky
    e_sun, e_sky = r.sun_and_sky_irradiance(point, np.array([1.0, 0.0, 0.0]), sun)
    assert np.allclose(e_sky, 0.1) and np.allclose(e_sun, 0.0)


def test_luminance_factors_scale_the_outputs(pas, orc):
    cp, o, r0, _, _ = uniform_renderer(pas, orc, False)
    sky_k, sun_k = np.array([2.0, 3.0, 5.0]), np.array([7.0, 11.0, 13.0])
    r1 = orc.Renderer(o, *[r0._keep[i] for i in (0, 1, 4)], single_mie=r0._keep[2], sky_k=sky_k, sun_k=sun_k)
    camera = np.array([0.0, 0.0, cp.bottom_radius + 1.0])
    view = np.array([0.0, math.sin(0.3), math.cos(0.3)])
    sun = np.array([math.sin(1.0), 0.0, math.cos(1.0)])
    assert np.allclose(r1.sky_radiance(camera, view, 0.0, sun)[0], r0.sky_radiance(camera, view, 0.0, sun)[0] * sky_k)
    a0, b0 = r0.sun_and_sky_irradiance(camera, view, sun)
    a1, b1 = r1.sun_and_sky_irradiance(camera, view, sun)
    assert np.allclose(a1, a0 * sun_k) and np.allclose(b1, b0 * sky_k)   # model.cc:272-280
    assert np.allclose(r1.solar_radiance(), r0.solar_radiance() * sun_k)
    alpha = cp.sun_angular_radius
    assert np.allclose(r0.solar_radiance(), cp.solar_irradiance / (math.pi * alpha * alpha))  # model.cc:228-231


def test_psnr_and_tone_map_follow_the_reference(pas):
    scn = scene
    rgb = np.zeros((2, 2, 3))
    rgb[0, 0] = [0.1, 0.2, 0.3]
    img = scn.tone_map(rgb, 10.0)
    want = [int((1.0 - math.exp(-v * 10.0)) ** (1 / 2.2) * 255.0) for v in (0.1, 0.2, 0.3)]
    assert img[0, 0] == (255 << 24) | (want[0] << 16) | (want[1] << 8) | want[2]
    assert img[1, 1] == 255 << 24
    other = img.copy()
    other[1, 1] = (255 << 24) | (10 << 16)
    # model_test.cc:750-765, as written: sqrt of the mean square error inside the log
    mse = math.sqrt(100.0 / 4)
    assert scn.psnr(img, other) == pytest.approx(10 * math.log10(255 * 255 / mse))
    assert scn.psnr(img, img) == float("inf")

"""GPU parity of the render-time path, through the C ABI:

  * the CUDA lookups (GetSkyRadiance, GetSkyRadianceToPoint, GetSunAndSkyIrradiance) and the
    model_test.glsl scene kernel against the fp64 oracle (oracle/pas_oracle_render.c, itself pinned
    to the unmodified reference by tests/test_oracle_render.py) evaluated on the GPU's OWN tables --
    this isolates the render kernels from the precompute;
  * GPU precompute + GPU render against the committed golden images of the UNMODIFIED reference
    (tests/golden/render_earth47.npz: 47-lane CPU model + model_test.glsl, oracle/run_reference_render.py):
    radiance within the 1e-3 contract per pixel, and the 13 test cases of
    atmosphere/reference/model_test.cc:805-1083 with the reference's own PSNR thresholds.

Tolerances: rendered radiance / luminance from the same tables, GPU (fp64 geometry, fp32 output)
vs oracle: 1e-5 relative (floored at 1e-6 of the image maximum without the sun disc). GPU tables
vs reference tables: the north_star's 1e-3 relative per pixel.
"""
import math
import os

import numpy as np
import pytest

from tests import scene, scene_render

from tests import parity

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(parity.GOLDEN, "render_earth47.npz")
LAM = [680.0, 550.0, 440.0]


def planar(tex):
    """RGBA texels [..., 4] -> planar float64 [4, ...]."""
    return np.ascontiguousarray(np.moveaxis(np.asarray(tex, dtype=np.float64), -1, 0))


def make_model(pas, n_wavelengths, combined, half, orders=4):
    spec = pas.model_test_earth(n_wavelengths, combine_scattering_textures=combined, half_precision=half)
    model = pas.Model.from_spec(spec)
    model.Init(orders)
    return spec, model


def oracle_renderer(pas, orc, spec, model, use_luminance):
    """fp64 oracle renderer over the tables the GPU model holds (read back as float32)."""
    cp = pas.channel_params(spec, LAM)
    o = orc.Oracle(cp)
    T = planar(model.transmittance)[:3]
    S4 = planar(model.scattering)
    E = planar(model.irradiance)[:3]
    k = model.luminance_factors()
    sky_k, sun_k = (k[0], k[1]) if use_luminance else (None, None)
    if model.combine_scattering_textures:
        return orc.Renderer(o, T, S4[:3], E, single_mie=None, scattering_alpha=S4[3:4], sky_k=sky_k, sun_k=sun_k)
    M = planar(model.single_mie_scattering)[:3]
    return orc.Renderer(o, T, S4[:3], E, single_mie=M, sky_k=sky_k, sun_k=sun_k)


def image_rel_error(got, want):
    """Per-pixel relative error, floored at 1e-6 of the brightest non-sun pixel."""
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    scale = np.percentile(np.abs(want), 99.0)
    return float(np.max(np.abs(got - want) / np.maximum(np.abs(want), 1e-6 * scale)))


@pytest.fixture(scope="module")
def models(pas):
    cache = {}

    def get(n_wavelengths, combined, half):
        key = (n_wavelengths, combined, half)
        if key not in cache:
            cache[key] = make_model(pas, n_wavelengths, combined, half)
        return cache[key]
    yield get
    for _, m in cache.values():
        m.close()


# ---- render kernels vs the oracle on the same tables ---------------------------------------------

@pytest.mark.parametrize("combined,half,use_luminance", [(False, False, False), (True, False, False),
                                                         (True, True, False), (False, True, True),
                                                         (True, False, True)])
@pytest.mark.parametrize("zenith", [65.0, 88.0, 96.0])
def test_scene_matches_oracle_on_the_same_tables(pas, orc, models, combined, half, use_luminance, zenith):
    spec, model = models(3, combined, half)
    view = scene.model_test_view(zenith, 90.0, use_luminance, width=160, height=90,
                                     sun_angular_radius=spec.sun_angular_radius)
    rgb, argb = scene_render.render_scene(model, view)
    want = oracle_renderer(pas, orc, spec, model, use_luminance).render_scene(view)
    assert np.isfinite(rgb).all()
    err = image_rel_error(rgb, want)
    assert err < 1e-5, err
    # tone-mapped words: identical up to the truncation of values sitting on an 8-bit boundary
    want_argb = scene.tone_map(want, view.exposure)
    ch = lambda a: np.stack([(a >> 16) & 255, (a >> 8) & 255, a & 255], -1).astype(int)
    diff = np.abs(ch(argb) - ch(want_argb))
    assert diff.max() <= 1 and (diff > 0).mean() < 1e-3
    assert (argb >> 24 == 255).all()


def random_queries(o, n, seed):
    rng = np.random.default_rng(seed)
    bottom, top = o.cp.bottom_radius, o.cp.top_radius
    unit = lambda v: v / np.linalg.norm(v, axis=-1, keepdims=True)
    d = unit(rng.normal(size=(n, 3)))
    radius = np.where(np.arange(n) % 3 == 2, top + rng.uniform(1.0, 3000.0, n), bottom + rng.uniform(1e-3, 59.0, n))
    camera = d * radius[:, None]
    view = unit(rng.normal(size=(n, 3)))
    outside = np.arange(n) % 3 == 2
    view[outside] = unit(-d[outside] + 0.02 * rng.normal(size=(int(outside.sum()), 3)))
    sun = unit(rng.normal(size=(n, 3)))
    shadow = np.where(np.arange(n) % 2 == 1, rng.uniform(0.0, 20.0, n), 0.0)
    return camera, view, sun, shadow, d


@pytest.mark.parametrize("combined,half", [(False, False), (True, True)])
def test_point_lookups_match_oracle(pas, orc, models, combined, half):
    spec, model = models(3, combined, half)
    r = oracle_renderer(pas, orc, spec, model, False)
    n = 600
    camera, view, sun, shadow, d = random_queries(r.o, n, 11)
    L, tr = model.GetSkyRadiance(camera, view, shadow, sun)
    scale = float(np.percentile(np.abs(L), 90))
    worst = 0.0
    for q in range(n):
        wl, wt = r.sky_radiance(camera[q], view[q], shadow[q], sun[q])
        if np.isnan(wl).any():
            continue
        worst = max(worst, float(np.max(np.abs(L[q] - wl) / np.maximum(np.abs(wl), 1e-5 * scale))),
                    float(np.max(np.abs(tr[q] - wt))))
    assert worst < 1e-5, worst
    # to-point: points along the view ray, cameras inside the atmosphere only
    inside = np.arange(n) % 3 != 2
    dist = np.random.default_rng(5).uniform(0.1, 200.0, n)
    point = camera + view * dist[:, None]
    ok = inside & (np.linalg.norm(point, axis=1) > r.o.cp.bottom_radius)
    sl = np.minimum(shadow, dist)
    L, tr = model.GetSkyRadianceToPoint(camera[ok], point[ok], sl[ok], sun[ok])
    worst = 0.0
    for i, q in enumerate(np.nonzero(ok)[0]):
        wl, wt = r.sky_radiance_to_point(camera[q], point[q], sl[q], sun[q])
        if np.isnan(wl).any() or np.isnan(wt).any():
            continue
        worst = max(worst, float(np.max(np.abs(L[i] - wl) / np.maximum(np.abs(wl), 1e-5 * scale))),
                    float(np.max(np.abs(tr[i] - wt))))
    assert worst < 1e-5, worst
    # sun and sky irradiance
    p = d * (r.o.cp.bottom_radius + np.random.default_rng(6).uniform(0.0, 60.0, n))[:, None]
    e_sun, e_sky = model.GetSunAndSkyIrradiance(p, view, sun)
    worst = 0.0
    for q in range(n):
        ws, wk = r.sun_and_sky_irradiance(p[q], view[q], sun[q])
        worst = max(worst, float(np.max(np.abs(e_sun[q] - ws) / np.maximum(np.abs(ws), 1e-6))),
                    float(np.max(np.abs(e_sky[q] - wk) / np.maximum(np.abs(wk), 1e-6))))
    assert worst < 1e-5, worst
    assert np.allclose(model.GetSolarRadiance(), r.solar_radiance(), rtol=1e-12)


def test_lookups_accept_device_pointers(pas, models):
    import ctypes
    import torch
    spec, model = models(3, True, False)
    n = 256
    r_cam = 6360.5
    camera = torch.tensor([[0.0, 0.0, r_cam]] * n, dtype=torch.float64, device="cuda")
    ang = torch.linspace(0.0, 1.5, n, dtype=torch.float64, device="cuda")
    view = torch.stack([torch.zeros_like(ang), torch.sin(ang), torch.cos(ang)], -1).contiguous()
    sun = torch.tensor([[0.0, math.sin(1.0), math.cos(1.0)]] * n, dtype=torch.float64, device="cuda")
    L = torch.zeros(n, 3, dtype=torch.float32, device="cuda")
    tr = torch.zeros(n, 3, dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    st = model._lib.pas_model_get_sky_radiance(model._h, 0, n, ctypes.c_void_p(camera.data_ptr()),
                                               ctypes.c_void_p(view.data_ptr()), None,
                                               ctypes.c_void_p(sun.data_ptr()), ctypes.c_void_p(L.data_ptr()),
                                               ctypes.c_void_p(tr.data_ptr()))
    assert st == 0
    host_L, host_tr = model.GetSkyRadiance(camera.cpu().numpy(), view.cpu().numpy(), None, sun.cpu().numpy())
    assert np.array_equal(L.cpu().numpy(), host_L) and np.array_equal(tr.cpu().numpy(), host_tr)
    assert (host_L > 0).all() and (host_tr > 0).all() and (host_tr <= 1).all()


def test_render_error_paths(pas, models):
    spec = pas.model_test_earth(3)
    fresh = pas.Model.from_spec(spec)
    view = scene.model_test_view(65.0, 90.0, False, width=16, height=9,
                                     sun_angular_radius=spec.sun_angular_radius)
    with pytest.raises(pas.PasError) as e:   # before Init
        scene_render.render_scene(fresh, view)
    assert e.value.status == 5
    fresh.close()
    # the radiance API does not exist with precomputed luminance (atmosphere/model.h:120-144)
    spec15, model15 = models(15, True, True)
    with pytest.raises(pas.PasError) as e:
        scene_render.render_scene(model15, view)
    assert e.value.status == 5
    with pytest.raises(pas.PasError):
        model15.GetSkyRadiance([0, 0, 6361.0], [0, 0, 1.0], None, [0, 0, 1.0])
    L, _ = model15.GetSkyRadiance([0, 0, 6361.0], [0, 0, 1.0], None, [0, 0, 1.0], use_luminance=True)
    assert np.isfinite(L).all() and (L > 0).all()
    bad = scene.model_test_view(65.0, 90.0, True, width=0, height=9,
                                    sun_angular_radius=spec.sun_angular_radius)
    with pytest.raises(ValueError):
        scene_render.render_scene(model15, bad)


# ---- GPU precompute + GPU render vs the reference's CPU images -------------------------------------

@pytest.fixture(scope="module")
def golden_images():
    return np.load(GOLDEN)


def golden_view(pas, spec, golden_images, zenith, use_luminance, constant_albedo=False):
    w, h = (int(v) for v in golden_images["size"])
    kw = dict(ground_albedo=[0.1] * 3, sphere_albedo=[0.8] * 3) if constant_albedo else {}
    return scene.model_test_view(zenith, 90.0, use_luminance, width=w, height=h,
                                     sun_angular_radius=spec.sun_angular_radius, **kw)


@pytest.mark.parametrize("zenith", [65, 88])
def test_radiance_image_within_contract_of_the_cpu_reference(pas, models, golden_images, zenith):
    """fp32 tables, separate Mie texture: the same computation as the CPU model at 680/550/440 nm
    (model_test.cc:797-803), so the north_star tolerance applies pixel by pixel."""
    spec, model = models(3, False, False)
    rgb, _ = scene_render.render_scene(model, golden_view(pas, spec, golden_images, zenith, False))
    want = golden_images[f"sun{zenith}_radiance"]
    err = np.abs(rgb.astype(np.float64) - want) / np.maximum(np.abs(want), 1e-6 * np.percentile(want, 99.0))
    assert np.isfinite(rgb).all()
    assert err.max() <= parity.REL_TOL, (err.max(), np.unravel_index(err.argmax(), err.shape))
    print(f"zenith {zenith}: max relative error vs the CPU reference image {err.max():.2e}")


# (name, wavelengths, combined, zenith, luminance, constant albedo, PSNR threshold) --
# atmosphere/reference/model_test.cc:805-1083
CASES = [
    ("RadianceSeparateTextures", 3, False, 65, False, False, 47.0),
    ("RadianceCombineTextures", 3, True, 65, False, False, 46.0),
    ("RadianceCombineTexturesSunSet", 3, True, 88, False, False, 40.0),
    ("LuminanceSeparateTexturesConstantAlbedo", 3, False, 65, True, True, 40.0),
    ("LuminanceCombineTexturesConstantAlbedo", 3, True, 65, True, True, 40.0),
    ("LuminanceCombineTexturesConstantAlbedoSunSet", 3, True, 88, True, True, 35.0),
    ("LuminanceCombineTexturesSpectralAlbedo", 3, True, 65, True, False, 38.0),
    ("LuminanceCombineTexturesSpectralAlbedoSunSet", 3, True, 88, True, False, 35.0),
    ("PrecomputedLuminanceSeparateTexturesConstantAlbedo", 15, False, 65, True, True, 43.0),
    ("PrecomputedLuminanceCombineTexturesConstantAlbedo", 15, True, 65, True, True, 43.0),
    ("PrecomputedLuminanceCombineTexturesConstantAlbedoSunSet", 15, True, 88, True, True, 40.0),
    ("PrecomputedLuminanceCombineTexturesSpectralAlbedo", 15, True, 65, True, False, 39.0),
    ("PrecomputedLuminanceCombineTexturesSpectralAlbedoSunSet", 15, True, 88, True, False, 40.0),
]


@pytest.mark.parametrize("name,n_wavelengths,combined,zenith,use_luminance,constant,threshold", CASES,
                         ids=[c[0] for c in CASES])
def test_reference_integration_cases_psnr(pas, models, golden_images, name, n_wavelengths, combined, zenith,
                                          use_luminance, constant, threshold):
    """The reference's own GPU-vs-CPU image tests: half-precision tables on the GPU side
    (model_test.cc:418), CPU image from the 47-lane model; PSNR as written in model_test.cc:750-765."""
    spec, model = models(n_wavelengths, combined, True)
    view = golden_view(pas, spec, golden_images, zenith, use_luminance, constant)
    _, argb = scene_render.render_scene(model, view)
    key = f"sun{zenith}_radiance" if not use_luminance else (
        f"sun{zenith}_luminance_{'constant' if constant else 'spectral'}")
    want = scene.tone_map(golden_images[key], view.exposure)
    psnr = scene.psnr(argb, want)
    print(f"{name}: PSNR {psnr:.1f} dB (reference threshold {threshold})")
    assert psnr > threshold, (name, psnr)

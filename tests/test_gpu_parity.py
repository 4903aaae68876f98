"""GPU parity tests: the CUDA path, called through the C ABI (include/pas_b200.h), against
  * the committed golden fixtures of the UNMODIFIED reference CPU model (tests/golden/, Earth/demo
    atmosphere at the reference's full table sizes, 15 spectral + RGB channels, 4 orders), and
  * the fp64 oracle (oracle/, checker only) on seeded small configurations, chained and
    teacher-forced, over the table sizes / flags / edge cases the reference supports.

Tolerance (north_star): max relative error <= 1e-3 per texel, fp32 GPU vs fp64 CPU. Texels whose
reference magnitude is below 1e-6 of the table's maximum are compared against that floor instead
(tests/parity.py). The measured errors are ~1e-5, so most assertions use a
tighter bound to catch regressions.
"""
import json
import os

import numpy as np
import pytest

from tests import parity

pytestmark = pytest.mark.gpu

TOL = parity.REL_TOL   # the contract
TIGHT = 1e-4           # what the kernels actually achieve, with margin


def assert_close(name, got, want, tol=TIGHT):
    m = parity.error_metrics(got, want)
    assert m["nan"] == 0, (name, m)
    assert m["max_floor"] <= tol and m["max_rel"] <= tol, (name, m)
    return m


def oracle_for(pas, orc, spec, model, sizes=None):
    cp = pas.channel_params(spec, model.channels())
    return orc.Oracle(cp, orc.Sizes(**(sizes or {})))


# ---- Earth at the reference's sizes vs the golden fixtures ----------------------------------------

@pytest.fixture(scope="module")
def earth_rgb(pas):
    spec = pas.earth(3, half_precision=False, max_sun_zenith_deg=102.0)
    model = pas.Model.from_spec(spec)
    model.set_capture(True)
    model.Init(4)
    yield spec, model
    model.close()


@pytest.fixture(scope="module")
def earth_spectral(pas):
    spec = pas.earth(15, half_precision=False, max_sun_zenith_deg=102.0)
    model = pas.Model.from_spec(spec)
    model.set_capture(True)
    model.Init(4)
    yield spec, model
    model.close()


NAMES_2D = ["transmittance"] + [f"delta_irradiance_{n}" for n in range(1, 5)]
NAMES_3D = ["delta_rayleigh", "delta_mie"] + [f"delta_{t}_{n}" for n in range(2, 5)
                                             for t in ("density", "multiple")]


@pytest.mark.parametrize("which", ["rgb", "spectral"])
def test_earth_intermediates_match_reference_golden(which, earth_rgb, earth_spectral, golden):
    two, three, _ = golden
    spec, model = earth_rgb if which == "rgb" else earth_spectral
    lanes = slice(15, 18) if which == "rgb" else slice(0, 15)
    report = {}
    for name in NAMES_2D:
        report[name] = assert_close(name, model.intermediate(name), two[name][lanes])
    for name in NAMES_3D:
        got = parity.sample3d(model.intermediate(name), three["indices"])
        report[name] = assert_close(name, got, three[name][lanes])
    worst = max(m["max_floor"] for m in report.values())
    assert worst <= TOL
    print(f"{which}: worst floored relative error over all tables {worst:.3e}")


@pytest.mark.parametrize("which", ["rgb", "spectral"])
def test_every_row_of_every_table_matches_the_reference_digest(which, earth_rgb, earth_spectral):
    """All 1,048,576 texels of every 3-D intermediate (not only the 3,240 sampled ones): sum, weighted
    sum and max of each of the 4096 rows of 256 texels against the digests of the reference run
    (tests/golden/earth18_rows.npz, oracle/gen_row_digest.py)."""
    rows = parity.load_rows()
    spec, model = earth_rgb if which == "rgb" else earth_spectral
    lanes = slice(15, 18) if which == "rgb" else slice(0, 15)
    worst = {}
    for name in NAMES_3D:
        m = parity.digest_metrics(model.intermediate(name), rows, name, lanes)
        assert m["nan"] == 0 and m["worst"] <= TIGHT, (name, m)
        worst[name] = m["worst"]
    if which == "rgb":
        # radiance mode: the product table is the per-channel sum of the reference (identity matrix)
        S = np.moveaxis(model.scattering[..., :3], -1, 0)
        m = parity.digest_metrics(S, rows, "scattering", lanes)
        assert m["worst"] <= TIGHT, m
        worst["scattering"] = m["worst"]
    print(f"{which}: worst row-digest error {max(worst.values()):.3e} over {len(worst)} tables x 4096 rows")


def test_bench_product_fp16_combined_luminance(pas, golden):
    """The product bench.py times (BASELINE config 2 with the demo's settings: 15 wavelengths,
    combined textures, HALF-precision final tables), against the reference: every row of S (fp16 RGBA)
    against the digests of the luminance table computed from the reference's fp64 tables, E and T
    against the golden 2-D tables. bench.py prints the same check as its `parity` key at every N."""
    two, _, _ = golden
    model = pas.Model.from_spec(pas.earth(15, half_precision=True, combine_scattering_textures=True))
    model.Init(4)
    assert model.texture_info(pas.TEXTURE_SCATTERING).bytes_per_channel == 2
    m = parity.check_bench_product(model.scattering, model.irradiance, model.transmittance,
                                   model.luminance_matrix(), two, parity.load_rows(), half_precision=True)
    assert m["ok"], m
    print("bench product parity:", m)
    model.close()


def test_earth_rgb_final_tables(earth_rgb, golden):
    two, three, _ = golden
    _, model = earth_rgb
    S = model.scattering                     # [R, MU, W, 4]
    got = np.moveaxis(S[..., :3], -1, 0)
    assert_close("scattering", parity.sample3d(got, three["indices"]), three["scattering"][15:18])
    # combined textures: alpha = red channel of single Mie (model.cc:154-155)
    alpha = parity.sample3d(S[..., 3][None], three["indices"])
    assert_close("scattering.a", alpha, three["delta_mie"][15:16])
    E = np.moveaxis(model.irradiance[..., :3], -1, 0)
    assert_close("irradiance", E, two["irradiance"][15:18])
    T = np.moveaxis(model.transmittance[..., :3], -1, 0)
    assert_close("transmittance", T, two["transmittance"][15:18], tol=1e-6)
    # BASELINE.md spot values (550 nm) are of a run with the test sun radius; T does not depend on it
    assert T[1, 32, 128] == pytest.approx(0.378841807, rel=1e-6)


def test_earth_spectral_luminance_tables(pas, orc, earth_spectral, golden):
    """Precomputed-luminance mode (model.cc:914-963): S.rgb = sum_b L_b (dR + sum_n dS_n / P_R(nu)),
    S.a = red of L.dM, E = sum L.dE_n, T at 680/550/440 nm. The oracle side is the host matvec of the
    reference's fp64 tables."""
    two, three, meta = golden
    spec, model = earth_spectral
    idx = three["indices"]
    L = model.luminance_matrix().astype(np.float64)
    gold = json.load(open(os.path.join(parity.GOLDEN, "luminance.json")))
    assert np.array_equal(model.luminance_matrix(),
                          np.asarray(gold["n15"]["luminance_from_radiance"], dtype=np.float32))
    o = oracle_for(pas, orc, spec, model)
    nu = np.array([o.rmumusnu_from_frag_coord(i + 0.5, j + 0.5, k + 0.5)[3] for k, j, i in idx])
    inter = {name: three[name][:15] for name in NAMES_3D}
    inter.update({name: two[name][:15] for name in NAMES_2D})
    inter["nu"] = nu
    S_want, A_want, _, E_want = orc.final_tables(inter, L, 4)
    S = model.scattering
    got = parity.sample3d(np.moveaxis(S[..., :3], -1, 0), idx)
    # the luminance matrix has negative entries (XYZ -> sRGB): compare against the table scale
    assert_close("scattering", got, S_want)
    assert_close("scattering.a", parity.sample3d(S[..., 3][None], idx), A_want[None])
    assert_close("irradiance", np.moveaxis(model.irradiance[..., :3], -1, 0), E_want)
    T = np.moveaxis(model.transmittance[..., :3], -1, 0)
    assert_close("transmittance", T, two["transmittance"][15:18], tol=1e-6)
    sky, sun = model.luminance_factors()
    assert sky == [683.0, 683.0, 683.0]            # model.cc:675-679
    assert np.allclose(sun, gold["sun_k"], rtol=1e-12)


def test_rgb_luminance_factors(earth_rgb):
    _, model = earth_rgb
    gold = json.load(open(os.path.join(parity.GOLDEN, "luminance.json")))
    sky, sun = model.luminance_factors()
    assert np.allclose(sky, gold["sky_k"], rtol=1e-12) and np.allclose(sun, gold["sun_k"], rtol=1e-12)


# ---- seeded small configurations vs the oracle ------------------------------------------------------

SMALL = [
    # (name, sizes, orders)
    ("default-ratio", dict(transmittance_width=64, transmittance_height=16, scattering_r=8,
                           scattering_mu=32, scattering_mu_s=8, scattering_nu=8,
                           irradiance_width=16, irradiance_height=4), 4),
    ("nu4", dict(transmittance_width=32, transmittance_height=8, scattering_r=4, scattering_mu=16,
                 scattering_mu_s=8, scattering_nu=4, irradiance_width=16, irradiance_height=4), 3),
    ("nu16-ragged", dict(transmittance_width=48, transmittance_height=12, scattering_r=5,
                         scattering_mu=14, scattering_mu_s=7, scattering_nu=16, irradiance_width=12,
                         irradiance_height=3), 3),
    ("minimal", dict(transmittance_width=4, transmittance_height=2, scattering_r=2, scattering_mu=4,
                     scattering_mu_s=2, scattering_nu=2, irradiance_width=2, irradiance_height=2), 3),
]


def oracle_sizes(sizes):
    m = dict(transmittance_width="t_w", transmittance_height="t_h", scattering_r="r",
             scattering_mu="mu", scattering_mu_s="mu_s", scattering_nu="nu",
             irradiance_width="e_w", irradiance_height="e_h")
    return {m[k]: v for k, v in sizes.items()}


def perturbed(pas, seed, n=3):
    """Earth with seeded random turbidity / ozone / albedo / sun size (the sweep axes of config 5)."""
    rng = np.random.default_rng(seed)
    spec = pas.earth(n, mie_scale_height=float(rng.uniform(800, 2500)),
                     ozone_dobson=float(rng.uniform(100, 500)),
                     ground_albedo=float(rng.uniform(0.0, 0.9)),
                     max_sun_zenith_deg=float(rng.uniform(95, 120)),
                     sun_angular_radius=float(rng.uniform(0.003, 0.02)))
    spec.mie_phase_function_g = float(rng.uniform(0.5, 0.9))
    return spec


@pytest.mark.parametrize("name,sizes,orders", SMALL, ids=[s[0] for s in SMALL])
@pytest.mark.parametrize("planet", ["small_planet", "earth-seed1"])
def test_chained_precompute_matches_oracle(pas, orc, name, sizes, orders, planet):
    spec = pas.small_planet() if planet == "small_planet" else perturbed(pas, 1)
    model = pas.Model.from_spec(spec, sizes=sizes)
    model.set_capture(True)
    model.Init(orders)
    want = oracle_for(pas, orc, spec, model, oracle_sizes(sizes)).precompute(orders)
    # coarse tables amplify the fp32 rounding of the table coordinates (a texel spans a larger
    # range of the integrand): the tiny configurations are held to the contract, not the tight bound
    tol = {"minimal": TOL, "nu16-ragged": 5e-4}.get(name, TIGHT)
    for key, ref in want.items():
        if key in ("nu", "scattering", "irradiance"):
            continue
        assert_close(f"{name}/{key}", model.intermediate(key), ref, tol=tol)
    assert_close("scattering", np.moveaxis(model.scattering[..., :3], -1, 0), want["scattering"], tol=tol)
    assert_close("irradiance", np.moveaxis(model.irradiance[..., :3], -1, 0), want["irradiance"], tol=tol)
    model.close()


@pytest.mark.parametrize("n", [6, 15, 21, 24])
def test_channel_counts(pas, orc, n):
    """Every channel-group size the kernels are instantiated for (6 -> 4+2, 15, 21 -> 16+4+1,
    24 -> 16+8 channels per launch), i.e. also the additive blending of the final tables across
    groups (model.cc:946-948)."""
    sizes = SMALL[1][1]
    spec = perturbed(pas, 2, n=n)
    model = pas.Model.from_spec(spec, sizes=sizes)
    assert len(model.channels()) == n
    assert np.allclose(model.channels(), pas.precomputed_wavelengths(n))
    model.set_capture(True)
    model.Init(3)
    want = oracle_for(pas, orc, spec, model, oracle_sizes(sizes)).precompute(3)
    for key in ("delta_rayleigh", "delta_density_2", "delta_irradiance_2", "delta_multiple_2",
                "delta_density_3", "delta_multiple_3"):
        assert_close(key, model.intermediate(key), want[key])
    L = model.luminance_matrix().astype(np.float64)
    S_want, A_want, _, E_want = orc.final_tables(want, L, 3)
    S = model.scattering
    assert_close("scattering", np.moveaxis(S[..., :3], -1, 0), S_want)
    assert_close("scattering.a", S[..., 3][None], A_want[None])
    assert_close("irradiance", np.moveaxis(model.irradiance[..., :3], -1, 0), E_want)
    model.close()


def test_teacher_forced_passes(pas, orc):
    """Each pass alone, fed with the ORACLE's inputs (cast to fp32), so that an error cannot hide
    behind or be blamed on the previous pass."""
    name, sizes, _ = SMALL[0]
    spec = perturbed(pas, 3)
    model = pas.Model.from_spec(spec, sizes=sizes)
    o = oracle_for(pas, orc, spec, model, oracle_sizes(sizes))
    w = o.precompute(3)
    model.run_phase("transmittance")
    assert_close("T", model.intermediate("transmittance"), w["transmittance"], tol=1e-6)
    model.write_intermediate("transmittance", w["transmittance"])
    model.run_phase("direct_irradiance")
    assert_close("dE1", model.intermediate("delta_irradiance"), w["delta_irradiance_1"], tol=1e-6)
    model.run_phase("single_scattering")
    assert_close("dR", model.intermediate("delta_rayleigh"), w["delta_rayleigh"])
    assert_close("dM", model.intermediate("delta_mie"), w["delta_mie"])
    model.write_intermediate("delta_rayleigh", w["delta_rayleigh"])
    model.write_intermediate("delta_mie", w["delta_mie"])
    model.write_intermediate("delta_irradiance", w["delta_irradiance_1"])
    model.run_phase("scattering_density", 2)
    assert_close("dJ2", model.intermediate("delta_density"), w["delta_density_2"])
    model.run_phase("indirect_irradiance", 1)
    assert_close("dE2", model.intermediate("delta_irradiance"), w["delta_irradiance_2"])
    model.write_intermediate("delta_density", w["delta_density_2"])
    model.run_phase("multiple_scattering", 2)
    assert_close("dS2", model.intermediate("delta_multiple"), w["delta_multiple_2"])
    model.write_intermediate("delta_multiple", w["delta_multiple_2"])
    model.write_intermediate("delta_irradiance", w["delta_irradiance_2"])
    model.run_phase("scattering_density", 3)
    assert_close("dJ3", model.intermediate("delta_density"), w["delta_density_3"])
    model.run_phase("indirect_irradiance", 2)
    assert_close("dE3", model.intermediate("delta_irradiance"), w["delta_irradiance_3"])
    model.close()


# ---- flags, formats and edge cases ---------------------------------------------------------------

def test_half_precision_and_separate_mie(pas):
    """half_precision packs only the final 3-D tables (model.cc:438-456); without combined textures
    the single-Mie table is its own RGB texture (model.cc:151-156, 757-763)."""
    sizes = SMALL[0][1]
    spec = pas.small_planet()
    spec.combine_scattering_textures = False
    full = pas.Model.from_spec(spec, sizes=sizes)
    full.Init(3)
    spec.half_precision = True
    half = pas.Model.from_spec(spec, sizes=sizes)
    half.Init(3)
    info = half.texture_info(pas.TEXTURE_SCATTERING)
    assert (info.bytes_per_channel, info.channels, info.present) == (2, 4, 1)
    assert full.texture_info(pas.TEXTURE_SCATTERING).bytes_per_channel == 4
    assert half.texture_info(pas.TEXTURE_SINGLE_MIE).present == 1
    for which in (pas.TEXTURE_SCATTERING, pas.TEXTURE_SINGLE_MIE):
        a, b = full.texture(which), half.texture(which)
        raw = half.texture(which, as_float32=False)
        assert raw.dtype == np.float16 and np.array_equal(raw.astype(np.float32), b)
        # fp16 rounding of the fp32 result (accumulated over 3 passes): 2^-11 relative per rounding
        scale = np.abs(a[..., :3]).max()
        assert np.abs(a[..., :3] - b[..., :3]).max() <= 3 * 2.0 ** -11 * scale
    # 2-D tables are never packed ("16F gives artifacts", model.cc:432)
    assert np.array_equal(full.transmittance, half.transmittance)
    assert np.array_equal(full.irradiance, half.irradiance)
    # separate Mie table == L . dM, and S.a is still written
    mie = full.single_mie_scattering
    S = full.scattering
    assert np.array_equal(mie[..., 0], S[..., 3])
    comb = pas.small_planet()
    model = pas.Model.from_spec(comb, sizes=sizes)
    model.Init(3)
    assert model.texture_info(pas.TEXTURE_SINGLE_MIE).present == 0
    with pytest.raises(pas.PasError):
        model.texture(pas.TEXTURE_SINGLE_MIE)
    assert np.array_equal(model.scattering, S)
    for m in (full, half, model):
        m.close()


def test_single_order_and_reinit(pas, orc):
    """Init(1) = single scattering only, E = 0 (model.cc:1137-1139, appendix D.1/D.5); Init may be
    called again with another order count and is deterministic."""
    sizes = SMALL[1][1]
    spec = pas.small_planet()
    model = pas.Model.from_spec(spec, sizes=sizes)
    with pytest.raises(pas.PasError):
        model.texture(pas.TEXTURE_SCATTERING)  # PAS_ERR_STATE before Init
    with pytest.raises(pas.PasError):
        model.Init(0)
    model.Init(1)
    want = oracle_for(pas, orc, spec, model, oracle_sizes(sizes)).precompute(1)
    S1 = model.scattering
    assert_close("S(1)", np.moveaxis(S1[..., :3], -1, 0), want["delta_rayleigh"])
    assert np.all(model.irradiance[..., :3] == 0)
    model.Init(3)
    S3 = model.scattering
    assert np.all(S3[..., :3] >= S1[..., :3])
    model.Init(3)
    assert np.array_equal(model.scattering, S3)      # idempotent, bit for bit
    model.Init(1)
    assert np.array_equal(model.scattering, S1)
    model.close()


def test_dat_export_and_device_pointers(pas, tmp_path):
    """.dat files = raw little-endian RGBA32F, x fastest (demo/webgl/precompute.cc:63-106)."""
    sizes = SMALL[1][1]
    model = pas.Model.from_spec(pas.small_planet(), sizes=sizes)
    model.Init(2)
    model.save_dat(str(tmp_path))
    files = sorted(os.listdir(tmp_path))
    assert files == ["irradiance.dat", "scattering.dat", "transmittance.dat"]
    for fn, tab in (("transmittance.dat", model.transmittance), ("scattering.dat", model.scattering),
                    ("irradiance.dat", model.irradiance)):
        raw = np.fromfile(os.path.join(tmp_path, fn), dtype="<f4")
        assert raw.size == tab.size and np.array_equal(raw.reshape(tab.shape), tab)
    # the whole WebGL hand-off (demo/webgl/precompute.cc:81-106): .dat files + the three shader texts
    web = tmp_path / "webgl"
    web.mkdir()
    glsl = tmp_path / "glsl"
    glsl.mkdir()
    (glsl / "definitions.glsl").write_text("// DEFINITIONS\n")
    (glsl / "functions.glsl").write_text("// FUNCTIONS\n")
    model.save_webgl(str(web), str(glsl), vertex_shader="// the demo's vertex shader\n",
                     fragment_shader="// the demo's fragment shader\n")
    assert sorted(os.listdir(web)) == ["atmosphere_shader.txt", "fragment_shader.txt", "irradiance.dat",
                                       "scattering.dat", "transmittance.dat", "vertex_shader.txt"]
    assert (web / "atmosphere_shader.txt").read_text() == model.GetShaderSource(str(glsl))
    assert (web / "atmosphere_shader.txt").read_text() == pas.shader_source(pas.small_planet(), str(glsl), sizes=sizes)
    assert (web / "vertex_shader.txt").read_text() == "// the demo's vertex shader\n"
    assert np.array_equal(np.fromfile(web / "scattering.dat", dtype="<f4").reshape(model.scattering.shape), model.scattering)
    model.save_webgl(str(web), str(glsl))          # the caller's shaders are optional
    assert model.device_ptr(pas.TEXTURE_SCATTERING) != 0
    with pytest.raises(AssertionError):
        model.texture(pas.TEXTURE_IRRADIANCE, out=np.empty(3, dtype=np.float32))
    model.close()


def test_shader_source(pas, tmp_path):
    """GetShaderSource = header (sizes, ATMOSPHERE constant, luminance constants) + definitions.glsl
    + functions.glsl + API wrappers (model.cc:691-744, 769-772). The two .glsl files are inputs read
    from the reference checkout; stand-ins are used here."""
    (tmp_path / "definitions.glsl").write_text("// DEFINITIONS\n")
    (tmp_path / "functions.glsl").write_text("// FUNCTIONS\n")
    model = pas.Model.from_spec(pas.earth(3, half_precision=True))
    src = model.GetShaderSource(str(tmp_path))
    assert src.startswith("#version 330")
    for needle in ("const int TRANSMITTANCE_TEXTURE_WIDTH = 256;", "const int SCATTERING_TEXTURE_NU_SIZE = 8;",
                   "#define COMBINED_SCATTERING_TEXTURES", "// DEFINITIONS", "// FUNCTIONS",
                   "const AtmosphereParameters ATMOSPHERE = AtmosphereParameters(",
                   "SKY_SPECTRAL_RADIANCE_TO_LUMINANCE", "#define RADIANCE_API_ENABLED",
                   "GetSkyRadiance(", "GetSunAndSkyIlluminance(", "uniform sampler3D scattering_texture;"):
        assert needle in src, needle
    assert src.index("// DEFINITIONS") < src.index("ATMOSPHERE =") < src.index("// FUNCTIONS")
    assert "6360" in src and "6420" in src
    lum = pas.Model.from_spec(pas.earth(15, half_precision=True))
    assert "#define RADIANCE_API_ENABLED" not in lum.GetShaderSource(str(tmp_path))  # model.cc:771
    with pytest.raises(pas.PasError):
        model.GetShaderSource(str(tmp_path / "missing"))
    model.close()
    lum.close()


def test_unsupported_sizes_are_rejected(pas):
    with pytest.raises(pas.PasError) as e:
        pas.Model.from_spec(pas.earth(3), sizes=dict(scattering_nu=32))
    assert e.value.status == 4
    with pytest.raises(pas.PasError) as e:
        pas.Model.from_spec(pas.earth(3), sizes=dict(scattering_mu=7))
    assert e.value.status == 1


# ---- size-independent properties at the full BASELINE sizes ------------------------------------------

def test_linearity_in_solar_irradiance(pas, earth_spectral):
    """Every table is linear in the solar irradiance; a factor 2 is exact in fp32, so the scaled
    run must reproduce the baseline bit for bit (all 15 channels, 4 orders, full sizes)."""
    spec, model = earth_spectral
    scaled = pas.earth(15, half_precision=False, max_sun_zenith_deg=102.0)
    scaled.solar_irradiance = [2.0 * v for v in scaled.solar_irradiance]
    other = pas.Model.from_spec(scaled)
    other.Init(4)
    assert np.array_equal(other.scattering, 2.0 * model.scattering)
    E2, E1 = other.irradiance, model.irradiance
    assert np.array_equal(E2[..., :3], 2.0 * E1[..., :3])
    assert np.array_equal(other.transmittance, model.transmittance)
    other.close()


def test_orders_converge_and_tables_are_sane(earth_rgb):
    _, model = earth_rgb
    prev = None
    for n in range(2, 5):
        dS = model.intermediate(f"delta_multiple_{n}")
        assert np.isfinite(dS).all() and (dS >= 0).all()
        if prev is not None:
            assert dS.sum() < prev          # each order carries less energy
        prev = dS.sum()
    T = model.intermediate("transmittance")
    assert (T > 0).all() and (T <= 1).all()
    S = model.scattering
    assert np.isfinite(S).all() and (S >= 0).all()


def test_ten_orders(pas, earth_rgb):
    """BASELINE config 3: 10 scattering orders. The first 4 orders are unaffected by asking for
    more (bit-identical intermediates) and the series converges."""
    spec, model4 = earth_rgb
    model = pas.Model.from_spec(spec)
    model.set_capture(True)
    model.Init(10)
    for name in ("delta_multiple_2", "delta_density_4", "delta_multiple_4", "delta_irradiance_4"):
        assert np.array_equal(model.intermediate(name), model4.intermediate(name)), name
    sums = [float(model.intermediate(f"delta_multiple_{n}").sum()) for n in range(2, 11)]
    assert all(b < a for a, b in zip(sums, sums[1:])) and sums[-1] < 1e-2 * sums[0]
    S10, S4 = model.scattering, model4.scattering
    assert (S10[..., :3] >= S4[..., :3]).all()
    assert model.last_launch_count() > 0
    model.close()


@pytest.mark.parametrize("which,orders", [("rgb", 5), ("spectral", 4)])
def test_overlapped_schedule_is_bit_identical(pas, which, orders):
    """Init without captures runs the irradiance passes and the final RGB transmittance on a side
    stream beside the multiple-scattering passes and alternates two multiple-scattering buffers;
    with captures every pass is enqueued on one stream in the reference's order
    (model.cc:1048-1215). Same kernels, same inputs: the products must not differ by one bit."""
    spec = pas.earth(3 if which == "rgb" else 15, half_precision=False, max_sun_zenith_deg=102.0)
    seq = pas.Model.from_spec(spec)
    seq.set_capture(True)
    seq.Init(orders)
    ovl = pas.Model.from_spec(spec)
    for _ in range(2):          # twice: re-Init reuses the rotated buffers
        ovl.Init(orders)
        assert np.array_equal(ovl.scattering, seq.scattering)
        assert np.array_equal(ovl.irradiance, seq.irradiance)
        assert np.array_equal(ovl.transmittance, seq.transmittance)
    seq.close()
    ovl.close()


@pytest.mark.parametrize("n,combined,half,orders", [(15, True, True, 4), (3, False, False, 3), (3, True, True, 1),
                                                   (24, True, False, 2)])
def test_registered_host_outputs_equal_blocking_reads(pas, n, combined, half, orders):
    """pas_model_set_host_outputs: Init fills registered host buffers with T / S / single Mie / E while
    it runs (S: written by the last multiple-scattering pass itself into a pinned destination, copied
    in bands behind the split pass into a pageable one). Same bytes as reading the tables after a plain
    Init -- also with captures on (plain copies at the end) and on re-Init."""
    import torch
    spec = pas.earth(n, half_precision=half, combine_scattering_textures=combined, max_sun_zenith_deg=102.0)
    ref = pas.Model.from_spec(spec)
    ref.Init(orders)
    which = [pas.TEXTURE_TRANSMITTANCE, pas.TEXTURE_SCATTERING, pas.TEXTURE_IRRADIANCE]
    if not combined:
        which.append(pas.TEXTURE_SINGLE_MIE)
    want = {w: ref.texture(w, as_float32=False) for w in which}
    for capture, pinned in ((False, True), (True, True), (False, False)):
        m = pas.Model.from_spec(spec)
        m.set_capture(capture)
        bufs = {w: torch.zeros(want[w].shape, dtype=torch.float16 if want[w].dtype == np.float16 else torch.float32)
                for w in which}
        bufs = {w: (t.pin_memory() if pinned else t).numpy() for w, t in bufs.items()}
        m.set_host_outputs(transmittance=bufs[pas.TEXTURE_TRANSMITTANCE], scattering=bufs[pas.TEXTURE_SCATTERING],
                           irradiance=bufs[pas.TEXTURE_IRRADIANCE],
                           single_mie_scattering=bufs.get(pas.TEXTURE_SINGLE_MIE))
        for _ in range(2):
            for b in bufs.values():
                b[...] = 0
            m.Init(orders)
            for w in which:
                assert np.array_equal(bufs[w], want[w]), (w, capture, pinned)
        m.close()
    ref.close()

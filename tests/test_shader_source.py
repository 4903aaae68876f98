"""GetShaderSource (pas_shader_source / pas_model_shader_source) against what the reference's
`glsl_header_factory_` and `kAtmosphereShader` specify (atmosphere/model.cc:221-281, 691-744, 769-772).
CPU only: the GLSL source is a function of the constructor parameters alone (pas_shader_source needs
no device). With /root/reference present the real definitions.glsl / functions.glsl and the API
prototypes of kAtmosphereShader are read from the checkout at test time (nothing is copied); without it
(the GPU box) stand-in .glsl files are used and the reference-dependent checks are skipped.

Deliberate deviation: the reference prints every constant with std::to_string,
i.e. 6 decimals (model.cc:636-652); this library prints 9 significant digits, because its tables are
computed from the exact values. The header must therefore agree with the reference's LINE BY LINE and
TOKEN BY TOKEN, every number within half a unit of the reference's 6th decimal (or of this library's
9th significant digit, for the large luminance factors)."""
import math
import os
import re

import numpy as np
import pytest

import precomputed_atmospheric_scattering_b200 as pas

REF = os.environ.get("PAS_REFERENCE", "/root/reference")
HAVE_REF = os.path.exists(os.path.join(REF, "atmosphere", "functions.glsl"))
NUMBER = re.compile(r"(?<![\w.])[-+]?(?:\d+\.\d*|\.\d+|\d+)(?:[eE][-+]?\d+)?(?![\w.])")


@pytest.fixture(scope="module")
def glsl_dir(tmp_path_factory):
    if HAVE_REF:
        return os.path.join(REF, "atmosphere")
    d = tmp_path_factory.mktemp("glsl")
    (d / "definitions.glsl").write_text("// DEFINITIONS\n")
    (d / "functions.glsl").write_text("// FUNCTIONS\n")
    return str(d)


def to_string6(v: float) -> str:
    """std::to_string(double): "%f"."""
    return "%f" % v


def reference_header(spec, definitions: str, functions: str, sizes=None) -> str:
    """The text `glsl_header_factory_({kLambdaR, kLambdaG, kLambdaB})` assembles (model.cc:691-744),
    restated from its specification: same lines, same order, std::to_string numbers."""
    lam = [680.0, 550.0, 440.0]
    u = spec.length_unit_in_meters
    interp = lambda v, scale: "vec3(" + ",".join(
        to_string6(pas.atmospheres.interpolate(spec.wavelengths, v, l) * scale) for l in lam) + ")"

    def profile(layers):
        layers = list(layers)
        while len(layers) < 2:
            layers.insert(0, pas.DensityProfileLayer(0.0, 0.0, 0.0, 0.0, 0.0))
        body = ",".join("DensityProfileLayer(" + ",".join(to_string6(x) for x in (
            l.width / u, l.exp_term, l.exp_scale * u, l.linear_term * u, l.constant_term)) + ")" for l in layers)
        return "DensityProfile(DensityProfileLayer[2](" + body + "))"

    sz = dict(transmittance_width=256, transmittance_height=64, scattering_r=32, scattering_mu=128,
              scattering_mu_s=32, scattering_nu=8, irradiance_width=64, irradiance_height=16)
    sz.update(sizes or {})
    gold = __import__("json").load(open(os.path.join(os.path.dirname(__file__), "golden", "luminance.json")))
    sky = [683.0] * 3 if spec.num_precomputed_wavelengths > 3 else gold["sky_k"]
    sun = gold["sun_k"]
    lines = ["#version 330", "#define IN(x) const in x", "#define OUT(x) out x", "#define TEMPLATE(x)",
             "#define TEMPLATE_ARGUMENT(x)", "#define assert(x)"]
    for name, key in (("TRANSMITTANCE_TEXTURE_WIDTH", "transmittance_width"),
                      ("TRANSMITTANCE_TEXTURE_HEIGHT", "transmittance_height"),
                      ("SCATTERING_TEXTURE_R_SIZE", "scattering_r"), ("SCATTERING_TEXTURE_MU_SIZE", "scattering_mu"),
                      ("SCATTERING_TEXTURE_MU_S_SIZE", "scattering_mu_s"), ("SCATTERING_TEXTURE_NU_SIZE", "scattering_nu"),
                      ("IRRADIANCE_TEXTURE_WIDTH", "irradiance_width"), ("IRRADIANCE_TEXTURE_HEIGHT", "irradiance_height")):
        lines.append(f"const int {name} = {sz[key]};")
    head = "\n".join(lines) + "\n" + ("#define COMBINED_SCATTERING_TEXTURES\n" if spec.combine_scattering_textures else "")
    atm = ("const AtmosphereParameters ATMOSPHERE = AtmosphereParameters(\n" +
           interp(spec.solar_irradiance, 1.0) + ",\n" + to_string6(spec.sun_angular_radius) + ",\n" +
           to_string6(spec.bottom_radius / u) + ",\n" + to_string6(spec.top_radius / u) + ",\n" +
           profile(spec.rayleigh_density) + ",\n" + interp(spec.rayleigh_scattering, u) + ",\n" +
           profile(spec.mie_density) + ",\n" + interp(spec.mie_scattering, u) + ",\n" +
           interp(spec.mie_extinction, u) + ",\n" + to_string6(spec.mie_phase_function_g) + ",\n" +
           profile(spec.absorption_density) + ",\n" + interp(spec.absorption_extinction, u) + ",\n" +
           interp(spec.ground_albedo, 1.0) + ",\n" + to_string6(math.cos(spec.max_sun_zenith_angle)) + ");\n" +
           "const vec3 SKY_SPECTRAL_RADIANCE_TO_LUMINANCE = vec3(" + ",".join(to_string6(k) for k in sky) + ");\n" +
           "const vec3 SUN_SPECTRAL_RADIANCE_TO_LUMINANCE = vec3(" + ",".join(to_string6(k) for k in sun) + ");\n")
    return head + definitions + atm + functions


def assert_same_up_to_number_format(ours: str, theirs: str):
    """Same text once every number is replaced by a placeholder; every number equal to the reference's
    within half a unit of its last printed decimal (or exactly, for integers)."""
    strip = lambda t: NUMBER.sub("#", t)
    a, b = strip(ours), strip(theirs)
    if a != b:
        la, lb = a.splitlines(), b.splitlines()
        for i, (x, y) in enumerate(zip(la, lb)):
            assert x == y, f"line {i}: ours {x!r} != reference {y!r}"
        assert len(la) == len(lb)
    na, nb = NUMBER.findall(ours), NUMBER.findall(theirs)
    assert len(na) == len(nb)
    for x, y in zip(na, nb):
        if re.fullmatch(r"[-+]?\d+", y):
            assert float(x) == float(y), (x, y)
        else:
            decimals = len(y.split(".")[1]) if "." in y and "e" not in y.lower() else 6
            # ours: 9 significant digits (beyond a GLSL float); the reference's: 6 decimals
            assert abs(float(x) - float(y)) <= max(0.5000001 * 10.0 ** -decimals, 5.1e-9 * abs(float(y))), (x, y)


@pytest.mark.parametrize("n,combined,half", [(3, True, True), (3, False, False), (15, True, True)])
def test_header_matches_the_reference_header_factory(glsl_dir, n, combined, half):
    spec = pas.earth(n, half_precision=half, combine_scattering_textures=combined)
    src = pas.shader_source(spec, glsl_dir)
    definitions = open(os.path.join(glsl_dir, "definitions.glsl")).read()
    functions = open(os.path.join(glsl_dir, "functions.glsl")).read()
    want = reference_header(spec, definitions, functions)
    assert len(src) > len(want)
    assert_same_up_to_number_format(src[:len(src) - len(src) + src.index(functions) + len(functions)], want)
    # the two GLSL files are included verbatim, in the reference's order
    assert src.count(definitions) == 1 and src.count(functions) == 1
    assert src.index(definitions) < src.index("const AtmosphereParameters ATMOSPHERE") < src.index(functions)
    tail = src[src.index(functions) + len(functions):]
    # model.cc:769-772: the radiance API only without precomputed illuminance
    assert tail.startswith("#define RADIANCE_API_ENABLED\n") == (n <= 3)
    for sampler in ("uniform sampler2D transmittance_texture;", "uniform sampler3D scattering_texture;",
                    "uniform sampler3D single_mie_scattering_texture;", "uniform sampler2D irradiance_texture;"):
        assert tail.count(sampler) == 1


def test_constants_parse_back_to_the_channel_parameters(glsl_dir):
    """The ATMOSPHERE constant carries the exact parameters the tables are computed from, to float
    precision (the reference truncates them to 6 decimals)."""
    spec = pas.earth(3, half_precision=True)
    cp = pas.channel_params(spec, [680.0, 550.0, 440.0])
    src = pas.shader_source(spec, glsl_dir)
    body = src[src.index("const AtmosphereParameters ATMOSPHERE"):src.index("const vec3 SKY_SPECTRAL")]
    vec3s = [np.array([float(v) for v in m.split(",")]) for m in re.findall(r"vec3\(([^)]*)\)", body)]
    names = ["solar_irradiance", "rayleigh_scattering", "mie_scattering", "mie_extinction",
             "absorption_extinction", "ground_albedo"]
    assert len(vec3s) == len(names)
    for got, name in zip(vec3s, names):
        assert np.allclose(got, np.asarray(getattr(cp, name), dtype=np.float64), rtol=2e-8, atol=0), name
    scalars = [float(v) for v in re.findall(r"^([-+0-9.eE]+),?$", body, flags=re.M)]
    assert scalars[:3] == pytest.approx([cp.sun_angular_radius, cp.bottom_radius, cp.top_radius], rel=1e-8)
    assert float(re.search(r"\n([-+0-9.eE]+)\);\n", body).group(1)) == pytest.approx(cp.mu_s_min, rel=1e-8)
    layers = [[float(x) for x in m.split(",")] for m in re.findall(r"DensityProfileLayer\(([-+0-9.eE,]+)\)", body)]
    assert np.allclose(np.asarray(layers).reshape(3, 2, 5), np.asarray(cp.profiles).reshape(3, 2, 5), rtol=2e-8, atol=0)


@pytest.mark.skipif(not HAVE_REF, reason="needs the reference checkout")
def test_api_wrappers_have_the_prototypes_of_the_reference_shader(glsl_dir):
    """kAtmosphereShader (model.cc:221-281), read from the checkout: same functions, same return types,
    same parameter lists, the radiance ones inside the RADIANCE_API_ENABLED block; and each wrapper
    forwards to the functions.glsl function the reference's forwards to, with the same scale factors."""
    text = open(os.path.join(REF, "atmosphere", "model.cc")).read()
    ref = text[text.index('const char kAtmosphereShader[] = R"(') + 36:]
    ref = ref[:ref.index(')";')]
    src = pas.shader_source(pas.earth(3, half_precision=True), glsl_dir)
    functions = open(os.path.join(glsl_dir, "functions.glsl")).read()
    ours = src[src.index(functions) + len(functions):]
    proto = re.compile(r"(\w+)\s+(Get\w+)\s*\(([^)]*)\)\s*\{")
    norm = lambda p: re.sub(r"\s+", " ", p.strip())
    protos = lambda t: [(r, n, norm(p)) for r, n, p in proto.findall(t)]
    assert protos(ours) == protos(ref) and len(protos(ref)) == 8

    def split(t):
        a, b = t.index("#ifdef RADIANCE_API_ENABLED"), t.index("#endif")
        return t[a:b], t[b:]

    for mine, theirs in zip(split(ours), split(ref)):
        assert [n for _, n, _ in protos(mine)] == [n for _, n, _ in protos(theirs)]
    # bodies: the callee and the luminance factors applied
    def bodies(t):
        out = {}
        for m in proto.finditer(t):
            end = t.index("\n}", m.end()) if "\n}" in t[m.end():] else len(t)
            nxt = proto.search(t, m.end())
            body = t[m.end():nxt.start() if nxt else len(t)]
            calls = re.findall(r"\b(Get\w+)\s*\(\s*ATMOSPHERE", body)
            out[m.group(2)] = (calls, sorted(set(re.findall(r"\b(S[UK][NY]_SPECTRAL_RADIANCE_TO_LUMINANCE)\b", body))),
                               "solar_irradiance" in body)
        return out
    assert bodies(ours) == bodies(ref)


def test_sizes_and_errors(glsl_dir, tmp_path):
    spec = pas.earth(3)
    src = pas.shader_source(spec, glsl_dir, sizes=dict(scattering_nu=16, scattering_mu_s=64, transmittance_width=512))
    for needle in ("const int SCATTERING_TEXTURE_NU_SIZE = 16;", "const int SCATTERING_TEXTURE_MU_S_SIZE = 64;",
                   "const int TRANSMITTANCE_TEXTURE_WIDTH = 512;", "const int SCATTERING_TEXTURE_R_SIZE = 32;"):
        assert needle in src
    with pytest.raises(pas.PasError) as e:
        pas.shader_source(spec, str(tmp_path / "missing"))
    assert e.value.status == 6           # PAS_ERR_IO
    spec.wavelengths = list(reversed(spec.wavelengths))
    with pytest.raises(pas.PasError) as e:
        pas.shader_source(spec, glsl_dir)
    assert e.value.status == 1           # PAS_ERR_INVALID_ARGUMENT

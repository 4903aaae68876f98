"""Device-side known-answer tests: the reference's analytic unit tests of the three integrals
(atmosphere/reference/functions_test.cc:1155-1189 scattering density, :1202-1226 multiple scattering,
:1325-1337 indirect irradiance) run against the CUDA KERNELS themselves -- constant input tables
written into the device buffers (pas_model_write_intermediate), one pass (pas_model_run_phase), and
the whole output table against the closed form, on the synthetic planet of the reference's unit tests
(functions_test.cc:51-61) with reduced table sizes. These test the kernels against mathematics, not
against another program; the reference's tolerance (kEpsilon = 1e-3, doubled or x10 where it does) is
kept."""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

EPS = 1e-3
SIZES = dict(transmittance_width=64, transmittance_height=16, scattering_r=8, scattering_mu=32,
             scattering_mu_s=8, scattering_nu=8, irradiance_width=16, irradiance_height=4)


@pytest.fixture(scope="module")
def planet(pas, orc):
    from tests.test_gpu_parity import oracle_for, oracle_sizes
    spec = pas.small_planet()
    model = pas.Model.from_spec(spec, sizes=SIZES)
    model.run_phase("transmittance")     # sizes the buffers; its table is replaced below
    o = oracle_for(pas, orc, spec, model, oracle_sizes(SIZES))
    cp = pas.channel_params(spec, model.channels())
    params = o.texel_params()            # [R, MU, W, 5] = (r, mu, mu_s, nu, hit) of every texel
    yield spec, model, cp, params
    model.close()


def shapes(model):
    T = model.intermediate("transmittance")
    return T.shape, (T.shape[0], SIZES["scattering_r"], SIZES["scattering_mu"],
                     SIZES["scattering_nu"] * SIZES["scattering_mu_s"]), (T.shape[0], SIZES["irradiance_height"],
                                                                         SIZES["irradiance_width"])


def densities(cp, params):
    """beta_R rho_R(h) + beta_M rho_M(h) per channel and texel (exponential profiles of the planet)."""
    h = params[..., 0] - cp.bottom_radius
    P = np.asarray(cp.profiles, dtype=np.float64).reshape(3, 2, 5)
    rho = lambda p: np.clip(P[p, 1, 1] * np.exp(P[p, 1, 2] * h) + P[p, 1, 3] * h + P[p, 1, 4], 0.0, 1.0)
    bR, bM = np.asarray(cp.rayleigh_scattering), np.asarray(cp.mie_scattering)
    return bR[:, None, None, None] * rho(0)[None] + bM[:, None, None, None] * rho(1)[None]


def test_density_of_isotropic_radiance(planet):
    """Incident radiance the same in all directions, no ground term: scattered radiance density =
    (beta_R rho_R + beta_M rho_M) * L everywhere, because both phase functions integrate to 1 over the
    sphere (functions_test.cc:1155-1172; 2 kEpsilon: the 16 x 32 quadrature of the phase functions)."""
    spec, model, cp, params = planet
    t_shape, s_shape, e_shape = shapes(model)
    L = 13.0
    model.write_intermediate("transmittance", np.ones(t_shape, dtype=np.float32))
    model.run_phase("density_setup")
    model.write_intermediate("delta_multiple", np.full(s_shape, L, dtype=np.float32))
    model.write_intermediate("delta_irradiance", np.zeros(e_shape, dtype=np.float32))
    model.run_phase("scattering_density", 3)
    got = model.intermediate("delta_density").astype(np.float64)
    want = densities(cp, params) * L
    assert np.isfinite(got).all()
    # the reference's own ray: on the ground, looking horizontally (layer 0, last mu row), 2 kEpsilon
    j = SIZES["scattering_mu"] - 1
    assert np.abs(got[:, 0, j, :] / want[:, 0, j, :] - 1.0).max() < 2 * EPS
    # every other view direction: same closed form, within the error of the 16 x 32 quadrature of the
    # forward Mie lobe (g = 0.8), which the fp64 CPU model shows too (1.1e-2 for the nadir rows)
    assert np.abs(got / want - 1.0).max() < 1.5e-2
    # order 2 reads the single-scattering tables with the phase functions applied at lookup
    # (functions.glsl:995-1003): with dR = c / P_R ... not constant in nu, so the analytic case is the
    # one where only one table is lit and the other phase function drops out: dM = 0, dR = L gives
    # sum over directions of L * P_R(nu1) * (coef) -- still an integral of products of phase
    # functions; the closed form needs isotropy, which order 2 does not have. Covered by the oracle.


def test_density_of_uniform_ground_irradiance(planet):
    """No incident radiance, uniform ground irradiance E, transmittance 1: on the ground, looking
    horizontally, the scattered density is (beta_R + beta_M) * albedo / (2 pi) * E -- the Lambertian
    ground fills the lower hemisphere and the phase functions are symmetric (functions_test.cc:1174-1188).
    Layer 0 is r = bottom exactly and its last mu row is mu = 0 exactly; every (mu_s, nu) of that row."""
    spec, model, cp, params = planet
    t_shape, s_shape, e_shape = shapes(model)
    E = 13.0
    model.write_intermediate("transmittance", np.ones(t_shape, dtype=np.float32))
    model.run_phase("density_setup")
    model.write_intermediate("delta_multiple", np.zeros(s_shape, dtype=np.float32))
    model.write_intermediate("delta_irradiance", np.full(e_shape, E, dtype=np.float32))
    model.run_phase("scattering_density", 3)
    got = model.intermediate("delta_density").astype(np.float64)
    j = SIZES["scattering_mu"] - 1
    assert params[0, j, 0, 0] == pytest.approx(cp.bottom_radius, rel=1e-15) and abs(params[0, j, 0, 1]) < 1e-12
    bR, bM, albedo = (np.asarray(v) for v in (cp.rayleigh_scattering, cp.mie_scattering, cp.ground_albedo))
    want = (bR + bM) * albedo / (2.0 * math.pi) * E
    assert np.abs(got[:, 0, j, :] / want[:, None] - 1.0).max() < 2 * EPS
    # rays that cannot see the ground get nothing: the top layer looking up
    assert np.abs(got[:, -1, j, :]).max() < np.abs(got[:, 0, j, :]).max()


def test_multiple_scattering_of_uniform_density(planet):
    """Uniform scattered density J, transmittance 1: the ray march returns J times the distance to the
    nearest boundary, for EVERY texel (functions_test.cc:1202-1226 checks two rays): the trapezoid of a
    constant is exact and every interpolation weight set sums to 1."""
    spec, model, cp, params = planet
    t_shape, s_shape, _ = shapes(model)
    J = 0.17
    model.write_intermediate("transmittance", np.ones(t_shape, dtype=np.float32))
    model.write_intermediate("delta_density", np.full(s_shape, J, dtype=np.float32))
    model.run_phase("multiple_scattering", 2)
    got = model.intermediate("delta_multiple").astype(np.float64)
    r, mu, hit = params[..., 0], params[..., 1], params[..., 4] > 0.5
    b, t = cp.bottom_radius, cp.top_radius
    d_top = np.maximum(-r * mu + np.sqrt(np.maximum(r * r * (mu * mu - 1.0) + t * t, 0.0)), 0.0)
    d_bot = np.maximum(-r * mu - np.sqrt(np.maximum(r * r * (mu * mu - 1.0) + b * b, 0.0)), 0.0)
    want = J * np.where(hit, d_bot, d_top)
    scale = want.max()
    assert np.abs(got - want[None]).max() < EPS * scale
    live = want > 1e-3 * scale
    assert np.abs(got[:, live] / want[live][None] - 1.0).max() < EPS


def test_indirect_irradiance_of_isotropic_radiance(planet):
    """Sky radiance 1 everywhere: ground irradiance = pi for every (r, mu_s) (functions_test.cc:1325-1337,
    10 kEpsilon there; the 16 x 64 midpoint quadrature of cos(theta) sin(theta) is 4e-4 high)."""
    spec, model, cp, params = planet
    _, s_shape, _ = shapes(model)
    model.write_intermediate("delta_multiple", np.ones(s_shape, dtype=np.float32))
    model.run_phase("indirect_irradiance", 2)
    got = model.intermediate("delta_irradiance").astype(np.float64)
    assert np.abs(got - math.pi).max() < 10 * EPS
    assert np.abs(got / (math.pi * (1.0 + (math.pi / 32) ** 2 / 6.0)) - 1.0).max() < 1e-4

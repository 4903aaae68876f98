"""Analytic known-answer tests of the fp64 oracle (oracle/pas_oracle.c).

The reference ships no golden table data; what pins its hot path are the 26 analytic property
checks of atmosphere/reference/functions_test.cc (small synthetic planet, :51-61, tolerance 1e-3
unless noted). They are restated here against the oracle's point functions: same planet, same
closed forms, same tolerances. CPU only.
"""
import math

import numpy as np
import pytest

from precomputed_atmospheric_scattering_b200.atmospheres import ChannelParams

EPS = 1e-3                       # functions_test.cc:49
SOLAR = 123.0                    # W/m^2/nm          (:51-52)
BOTTOM, TOP = 1000.0, 1500.0     # km                (:53-54)
H_RAYLEIGH, H_MIE = 60.0, 30.0   # km                (:56-57)
K_RAYLEIGH, K_MIE_SCA, K_MIE_EXT = 0.001, 0.0015, 0.002  # 1/km (:58-60)
ALBEDO = 0.1


def planet(**kw) -> ChannelParams:
    """One-channel parameter block of the test planet (functions_test.cc:301-311); lengths in km."""
    exp_layer = lambda h: [[0, 0, 0, 0, 0], [0, 1.0, -1.0 / h, 0, 0]]
    p = dict(
        lambdas=np.array([550.0]), solar_irradiance=np.array([SOLAR]),
        rayleigh_scattering=np.array([K_RAYLEIGH]), mie_scattering=np.array([K_MIE_SCA]),
        mie_extinction=np.array([K_MIE_EXT]), absorption_extinction=np.array([0.0]),
        ground_albedo=np.array([ALBEDO]), sun_angular_radius=0.00935 / 2, bottom_radius=BOTTOM,
        top_radius=TOP, mie_phase_function_g=0.8, mu_s_min=-1.0,
        profiles=np.array([exp_layer(H_RAYLEIGH), exp_layer(H_MIE), [[0] * 5, [0] * 5]], dtype=float))
    p.update(kw)
    return ChannelParams(**p)


UNIFORM = [[0, 0, 0, 0, 0], [0, 0, 0, 0, 1.0]]  # SetUniformAtmosphere (:289-297): density == 1
NOTHING = [[0] * 5, [0] * 5]


def uniform_planet(aerosols=True):
    prof = np.array([UNIFORM, UNIFORM if aerosols else NOTHING, NOTHING], dtype=float)
    kw = dict(profiles=prof)
    if not aerosols:  # RemoveAerosols
        kw.update(mie_scattering=np.array([0.0]), mie_extinction=np.array([0.0]))
    return planet(**kw)


def horizon_mu(r):
    return -math.sqrt(1.0 - (BOTTOM / r) ** 2)


@pytest.fixture(scope="module")
def O(orc):
    return lambda cp=None, **sz: orc.Oracle(cp or planet(), orc.Sizes(**sz))


def test_distance_to_top(O):  # :321-333
    o, r = O(), BOTTOM * 0.2 + TOP * 0.8
    assert abs(o.distance_to_top(r, 1.0) - (TOP - r)) < 1e-3
    assert abs(o.distance_to_top(r, 0.0) - math.sqrt(TOP * TOP - r * r)) < 1e-3


def test_ray_intersects_ground(O):  # :343-352
    o, r = O(), BOTTOM * 0.9 + TOP * 0.1
    mh = horizon_mu(r)
    assert not o.ray_intersects_ground(r, 1.0)
    assert not o.ray_intersects_ground(r, mh + EPS)
    assert o.ray_intersects_ground(r, mh - EPS)
    assert o.ray_intersects_ground(r, -1.0)


def test_optical_length(O):  # :370-388
    r = BOTTOM * 0.2 + TOP * 0.8
    want = H_RAYLEIGH * (math.exp(-(r - BOTTOM) / H_RAYLEIGH) - math.exp(-(TOP - BOTTOM) / H_RAYLEIGH))
    assert abs(O().optical_length_to_top(0, r, 1.0) - want) < 1e-3
    got = O(uniform_planet()).optical_length_to_top(0, r, 0.0)
    assert abs(got - math.sqrt(TOP * TOP - r * r)) < 1e-3


def test_profile_density(O):  # :396-419
    one = lambda layer: planet(profiles=np.array([[[0] * 5, layer], NOTHING, NOTHING], dtype=float))
    o = O(one([0, 1.0, -1.0, 0, 0]))
    assert o.profile_density(0, 2.0) == math.exp(-2.0)
    o = O(one([0, 0, 0, -0.5, 1.0]))
    assert [o.profile_density(0, h) for h in (0.0, 1.0, 3.0)] == [1.0, 0.5, 0.0]
    tri = [[25.0, 0, 0, 1 / 15.0, -2 / 3.0], [0, 0, 0, -1 / 15.0, 8 / 3.0]]
    o = O(planet(profiles=np.array([tri, NOTHING, NOTHING], dtype=float)))
    assert o.profile_density(0, 0.0) == 0.0 and o.profile_density(0, 50.0) == 0.0
    for h, want in ((10.0, 0.0), (25.0, 1.0), (40.0, 0.0)):
        assert abs(o.profile_density(0, h) - want) < EPS


def test_transmittance_to_top(O):  # :433-469
    r = BOTTOM * 0.2 + TOP * 0.8
    h_r, h_top = r - BOTTOM, TOP - BOTTOM
    tau_r = K_RAYLEIGH * H_RAYLEIGH * (math.exp(-h_r / H_RAYLEIGH) - math.exp(-h_top / H_RAYLEIGH))
    tau_m = K_MIE_EXT * H_MIE * (math.exp(-h_r / H_MIE) - math.exp(-h_top / H_MIE))
    assert abs(O().compute_transmittance_to_top(r, 1.0)[0] - math.exp(-(tau_r + tau_m))) < EPS
    tri = [[25.0, 0, 0, 1 / 15.0, -2 / 3.0], [0, 0, 0, -1 / 15.0, 8 / 3.0]]
    ozone_only = planet(profiles=np.array([NOTHING, NOTHING, tri], dtype=float),
                        absorption_extinction=np.array([0.02]))
    assert abs(O(ozone_only).compute_transmittance_to_top(BOTTOM, 1.0)[0] - math.exp(-0.02 * 15.0)) < EPS
    got = O(uniform_planet(aerosols=False)).compute_transmittance_to_top(r, 0.0)[0]
    assert abs(got - math.exp(-K_RAYLEIGH * math.sqrt(TOP * TOP - r * r))) < EPS


def test_transmittance_uv_mapping(O):  # :479-574
    o = O()
    tw, th = 256, 64
    u, v = o.transmittance_uv_from_rmu(BOTTOM, 1.0)
    assert abs(u - 0.5 / tw) < EPS and abs(v - 0.5 / th) < EPS
    u, v = o.transmittance_uv_from_rmu(TOP, 1.0)
    assert abs(u - 0.5 / tw) < EPS and abs(v - (1 - 0.5 / th)) < EPS
    r, mu = o.rmu_from_transmittance_uv(0.5 / tw, 0.5 / th)
    assert abs(r - BOTTOM) < 1e-3 and abs(mu - 1.0) < EPS
    r, mu = o.rmu_from_transmittance_uv(1 - 0.5 / tw, 1 - 0.5 / th)
    assert abs(r - TOP) < 1e-3 and abs(mu - horizon_mu(TOP)) < EPS
    # round trip
    r0, mu0 = BOTTOM * 0.2 + TOP * 0.8, 0.25
    u, v = o.transmittance_uv_from_rmu(r0, mu0)
    r, mu = o.rmu_from_transmittance_uv(u, v)
    assert abs(r - r0) < 1e-3 and abs(mu - mu0) < EPS


def test_transmittance_lookup_equals_compute(O):  # :583-632
    o = O(uniform_planet(aerosols=False))
    T = o.transmittance()
    r, d = BOTTOM * 0.2 + TOP * 0.8, (TOP - BOTTOM) * 0.1
    want = math.exp(-K_RAYLEIGH * d)
    assert abs(o.get_transmittance(T, BOTTOM, 0.0, d, False)[0] - want) < EPS
    assert abs(o.get_transmittance(T, r, 0.7, d, False)[0] - want) < EPS
    assert abs(o.get_transmittance(T, r, -0.7, d, o.ray_intersects_ground(r, -0.7))[0] - want) < EPS


def _integrand(o, T, r, mu, mu_s, nu, d, hit):
    """ComputeSingleScatteringIntegrand (functions.glsl:650-668) from the oracle's point functions:
    T(p -> q) * T_sun(q) * density(q), with q at distance d along the ray."""
    r_d = min(max(math.sqrt(d * d + 2.0 * r * mu * d + r * r), BOTTOM), TOP)
    mu_s_d = min(max((r * mu_s + d * nu) / r_d, -1.0), 1.0)
    t = o.get_transmittance(T, r, mu, d, hit)[0] * o.get_transmittance_to_sun(T, r_d, mu_s_d)[0]
    return t * o.profile_density(0, r_d - BOTTOM), t * o.profile_density(1, r_d - BOTTOM)


def test_single_scattering_integrand(O):  # :676-736
    o = O()
    T = o.transmittance()
    h_top = TOP - BOTTOM
    h = h_top / 2.0
    # vertical ray from the ground, sun at the zenith, scattering in the middle of the atmosphere
    ray, mie = _integrand(o, T, BOTTOM, 1.0, 1.0, 1.0, h, False)
    tau_r = K_RAYLEIGH * H_RAYLEIGH * (1 - math.exp(-h_top / H_RAYLEIGH))
    tau_m = K_MIE_EXT * H_MIE * (1 - math.exp(-h_top / H_MIE))
    assert abs(ray - math.exp(-tau_r - tau_m) * math.exp(-h / H_RAYLEIGH)) < EPS
    assert abs(mie - math.exp(-tau_r - tau_m) * math.exp(-h / H_MIE)) < EPS
    # vertical ray looking down from the top boundary, scattering angle 180 degrees
    ray, mie = _integrand(o, T, TOP, -1.0, 1.0, -1.0, h, True)
    tau_r = 2 * K_RAYLEIGH * H_RAYLEIGH * (math.exp(-h / H_RAYLEIGH) - math.exp(-h_top / H_RAYLEIGH))
    tau_m = 2 * K_MIE_EXT * H_MIE * (math.exp(-h / H_MIE) - math.exp(-h_top / H_MIE))
    assert abs(ray - math.exp(-tau_r - tau_m) * math.exp(-h / H_RAYLEIGH)) < EPS
    assert abs(mie - math.exp(-tau_r - tau_m) * math.exp(-h / H_MIE)) < EPS
    # horizontal ray from the ground, sun at the horizon, uniform air without aerosols
    o = O(uniform_planet(aerosols=False))
    T = o.transmittance()
    ray, _ = _integrand(o, T, BOTTOM, 0.0, 0.0, 1.0, 50.0, False)
    assert abs(ray - math.exp(-K_RAYLEIGH * math.sqrt(TOP * TOP - BOTTOM * BOTTOM))) < EPS


def test_distance_to_nearest_boundary(O):  # :745-780; DistanceToNearestAtmosphereBoundary, functions.glsl:680-687
    o = O()
    nearest = lambda r, mu: (o.distance_to_bottom(r, mu) if o.ray_intersects_ground(r, mu)
                             else o.distance_to_top(r, mu))
    r = BOTTOM * 0.2 + TOP * 0.8
    assert abs(nearest(r, 1.0) - (TOP - r)) < 1e-3                        # 1 m
    assert abs(nearest(r, 0.0) - math.sqrt(TOP * TOP - r * r)) < 1e-3
    assert abs(nearest(r, -1.0) - (r - BOTTOM)) < 1e-3


def test_single_scattering_analytic(O):  # :818-858
    o = O()
    T = o.transmittance()
    h_top = TOP - BOTTOM
    ray, mie = o.single_scattering_point(T, BOTTOM, 1.0, 1.0, 1.0, False)
    tau_r = K_RAYLEIGH * H_RAYLEIGH * (1 - math.exp(-h_top / H_RAYLEIGH))
    tau_m = K_MIE_EXT * H_MIE * (1 - math.exp(-h_top / H_MIE))
    assert abs(ray[0] / (SOLAR * tau_r * math.exp(-tau_r - tau_m)) - 1) < 10 * EPS
    assert abs(mie[0] / (SOLAR * tau_m * K_MIE_SCA / K_MIE_EXT * math.exp(-tau_r - tau_m)) - 1) < 10 * EPS
    clear = planet(mie_scattering=np.array([0.0]), mie_extinction=np.array([0.0]))
    o = O(clear)
    T = o.transmittance()
    ray, mie = o.single_scattering_point(T, TOP, -1.0, 1.0, -1.0, True)
    want = SOLAR * 0.5 * (1 - math.exp(-2 * H_RAYLEIGH * K_RAYLEIGH * (1 - math.exp(-h_top / H_RAYLEIGH))))
    assert abs(ray[0] / want - 1) < 2 * EPS
    assert abs(mie[0]) < EPS


def test_phase_functions_integrate_to_one(orc):  # :865-879
    l = orc.lib()
    n = 100
    ray = mie = 0.0
    for i in range(n):
        theta = (i + 0.5) * math.pi / n
        dw = math.sin(theta) * (math.pi / n) * 2 * math.pi
        ray += l.paso_rayleigh_phase(ctypes_double(math.cos(theta))) * dw
        mie += l.paso_mie_phase(ctypes_double(0.8), ctypes_double(math.cos(theta))) * dw
    assert abs(ray - 1) < 2 * EPS and abs(mie - 1) < 2 * EPS


def ctypes_double(v):
    import ctypes
    return ctypes.c_double(v)


def test_scattering_uvwz_mapping(O):  # :888-1060
    o = O()
    NR, NMU, NMUS, NNU = 32, 128, 32, 8
    assert abs(o.scattering_uvwz_from_rmumusnu(BOTTOM, 0, 0, 0, False)[3] - 0.5 / NR) < EPS
    assert abs(o.scattering_uvwz_from_rmumusnu(TOP, 0, 0, 0, False)[3] - (1 - 0.5 / NR)) < EPS
    r = (TOP + BOTTOM) / 2
    mh = horizon_mu(r)
    assert abs(o.scattering_uvwz_from_rmumusnu(r, mh, 0, 0, True)[2] - 0.5 / NMU) < EPS
    assert abs(o.scattering_uvwz_from_rmumusnu(r, mh, 0, 0, False)[2] - (1 - 0.5 / NMU)) < EPS
    assert o.scattering_uvwz_from_rmumusnu(r, -1, 0, 0, True)[2] < 0.5
    assert o.scattering_uvwz_from_rmumusnu(r, 1, 0, 0, False)[2] > 0.5
    for rr in (BOTTOM, TOP):
        assert abs(o.scattering_uvwz_from_rmumusnu(rr, 0, -1, 0, False)[1] - 0.5 / NMUS) < EPS
        assert abs(o.scattering_uvwz_from_rmumusnu(rr, 0, 1, 0, False)[1] - (1 - 0.5 / NMUS)) < EPS
    assert abs(o.scattering_uvwz_from_rmumusnu(BOTTOM, 0, 0, -1, False)[0] - 0.0) < EPS
    assert abs(o.scattering_uvwz_from_rmumusnu(BOTTOM, 0, 0, 1, False)[0] - 1.0) < EPS


def test_scattering_inverse_mapping(O):  # :960-1060
    o = O()
    NR, NMU, NMUS = 32, 128, 32
    inv = o.rmumusnu_from_scattering_uvwz
    lo = [0.0, 0.5 / NMUS, 0.5 / NMU, 0.5 / NR]
    assert abs(inv(lo)[0] - BOTTOM) < 1e-3
    assert abs(inv([0.0, 0.5 / NMUS, 0.5 / NMU, 1 - 0.5 / NR])[0] - TOP) < 1e-3
    r, mu, _, _, hit = inv([0.0, 0.5 / NMUS, 0.5 / NMU + EPS, 0.5])
    assert abs(mu - horizon_mu(r)) < EPS and mu <= horizon_mu(r) and hit == 1.0
    r, mu, _, _, hit = inv([0.0, 0.5 / NMUS, 1 - 0.5 / NMU - EPS, 0.5])
    assert abs(mu - horizon_mu(r)) < 5 * EPS and mu >= horizon_mu(r) and hit == 0.0
    assert abs(inv(lo)[2] + 1.0) < EPS
    assert abs(inv([0.0, 1 - 0.5 / NMUS, 0.5 / NMU, 0.5 / NR])[2] - 1.0) < EPS
    assert abs(inv(lo)[3] + 1.0) < EPS
    assert abs(inv([1.0, 0.5 / NMUS, 0.5 / NMU, 0.5 / NR])[3] - 1.0) < EPS
    for rr in (BOTTOM, TOP):
        out = inv(o.scattering_uvwz_from_rmumusnu(rr, -1.0, 1.0, -1.0, True))
        assert abs(out[0] - rr) < 1e-3 and out[4] == 1.0
        assert np.allclose(out[1:4], [-1.0, 1.0, -1.0], atol=EPS)
    rm = (BOTTOM + TOP) / 2
    out = inv(o.scattering_uvwz_from_rmumusnu(rm, 0.2, 0.3, 0.4, False))
    assert abs(out[0] - rm) < 1e-3 and out[4] == 0.0
    assert np.allclose(out[1:4], [0.2, 0.3, 0.4], atol=EPS)


def test_frag_coord_nu_is_clamped_to_valid_range(O):  # functions.glsl:905-926
    o = O()
    for (x, y, z) in ((0.5, 0.5, 0.5), (255.5, 127.5, 31.5), (40.5, 64.5, 3.5), (200.5, 10.5, 20.5)):
        r, mu, mu_s, nu, hit = o.rmumusnu_from_frag_coord(x, y, z)
        s = math.sqrt(max((1 - mu * mu) * (1 - mu_s * mu_s), 0.0))
        assert mu * mu_s - s - 1e-12 <= nu <= mu * mu_s + s + 1e-12
        assert BOTTOM - 1e-9 <= r <= TOP + 1e-9 and hit == (1.0 if y < 64 else 0.0)


def test_scattering_density_analytic(O, orc):  # :1155-1189
    sz = dict(t_w=8, t_h=4, r=4, mu=8, mu_s=4, nu=4, e_w=4, e_h=2)
    o = O(**sz)
    full_T = np.ones((1, 4, 8))
    zeros = np.zeros((1, 4, 8, 16))
    uniform = np.full((1, 4, 8, 16), 13.0)
    no_E, uni_E = np.zeros((1, 2, 4)), np.full((1, 2, 4), 13.0)
    got = o.scattering_density_point(full_T, zeros, zeros, uniform, no_E, BOTTOM, 0.0, 0.0, 1.0, 3)[0]
    assert abs(got / ((K_RAYLEIGH + K_MIE_SCA) * 13.0) - 1) < 2 * EPS
    got = o.scattering_density_point(full_T, zeros, zeros, zeros, uni_E, BOTTOM, 0.0, 0.0, 1.0, 3)[0]
    assert abs(got / ((K_RAYLEIGH + K_MIE_SCA) * ALBEDO / (2 * math.pi) * 13.0) - 1) < 2 * EPS


def test_multiple_scattering_analytic(O):  # :1202-1226
    sz = dict(t_w=8, t_h=4, r=4, mu=8, mu_s=4, nu=4, e_w=4, e_h=2)
    o = O(**sz)
    full_T = np.ones((1, 4, 8))
    J = np.full((1, 4, 8, 16), 0.17)
    r = BOTTOM * 0.2 + TOP * 0.8
    got = o.multiple_scattering_point(full_T, J, r, -1.0, 1.0, -1.0, True)[0]
    assert abs(got - 0.17 * (r - BOTTOM)) < 0.17 * (r - BOTTOM) * EPS
    mu = horizon_mu(TOP)
    d = math.sqrt(TOP * TOP - BOTTOM * BOTTOM)
    got = o.multiple_scattering_point(full_T, J, TOP, mu, 1.0, mu, True)[0]
    assert abs(got - 0.17 * d) < 0.17 * d * EPS


def test_indirect_irradiance_analytic(O):  # :1325-1337
    sz = dict(t_w=8, t_h=4, r=4, mu=8, mu_s=4, nu=4, e_w=4, e_h=2)
    o = O(**sz)
    zeros, ones = np.zeros((1, 4, 8, 16)), np.ones((1, 4, 8, 16))
    got = o.indirect_irradiance_point(zeros, zeros, ones, BOTTOM, 1.0, 2)[0]
    assert abs(got - math.pi) < 10 * EPS


def test_irradiance_mapping(O):  # :1346-1399
    o = O()
    EW, EH = 64, 16
    assert abs(o.irradiance_uv_from_rmus(BOTTOM, 0.0)[1] - 0.5 / EH) < EPS
    assert abs(o.irradiance_uv_from_rmus(TOP, 0.0)[1] - (1 - 0.5 / EH)) < EPS
    assert abs(o.irradiance_uv_from_rmus(BOTTOM, -1.0)[0] - 0.5 / EW) < EPS
    assert abs(o.irradiance_uv_from_rmus(BOTTOM, 1.0)[0] - (1 - 0.5 / EW)) < EPS
    assert abs(o.rmus_from_irradiance_uv(0.5, 0.5 / EH)[0] - BOTTOM) < 1e-3
    assert abs(o.rmus_from_irradiance_uv(0.5, 1 - 0.5 / EH)[0] - TOP) < 1e-3
    assert abs(o.rmus_from_irradiance_uv(0.5 / EW, 0.5)[1] + 1.0) < EPS
    assert abs(o.rmus_from_irradiance_uv(1 - 0.5 / EW, 0.5)[1] - 1.0) < EPS


def test_lookup_equals_compute_through_tables(O):  # :1069-1131, 1237-1313, 1408-1443
    """Interpolated lookups in precomputed tables reproduce the direct computation at texel
    centres exactly and in between within the reference's tolerance."""
    sz = dict(t_w=64, t_h=16, r=8, mu=32, mu_s=8, nu=4, e_w=16, e_h=8)
    o = O(**sz)
    T = o.transmittance()
    dR, dM = o.single_scattering(T)
    for (x, y, z) in ((5.5, 20.5, 3.5), (17.5, 3.5, 6.5), (30.5, 31.5, 0.5)):
        r, mu, mu_s, nu, hit = o.rmumusnu_from_frag_coord(x, y, z)
        ray, mie = o.single_scattering_point(T, r, mu, mu_s, nu, hit)
        k, j, i = int(z), int(y), int(x)
        assert ray[0] == pytest.approx(dR[0, k, j, i], rel=1e-12, abs=1e-300)
        # the lookup at the texel's own coordinates returns the texel (nu is clamped per texel, so
        # only texels whose nu was not clamped sit exactly on a slab)
        nu_slab = -1.0 + 2.0 * (i // 8) / 3.0
        if abs(nu - nu_slab) < 1e-12:
            assert o.get_scattering(dR, r, mu, mu_s, nu, hit)[0] == pytest.approx(dR[0, k, j, i], rel=1e-9)
            assert o.get_scattering(dM, r, mu, mu_s, nu, hit)[0] == pytest.approx(mie[0], rel=1e-9)
    dE = o.direct_irradiance(T)
    r, mu_s = o.rmus_from_irradiance_uv(3.5 / 16, 2.5 / 8)
    assert o.get_irradiance(dE, r, mu_s)[0] == pytest.approx(o.direct_irradiance_point(T, r, mu_s)[0], rel=1e-9)

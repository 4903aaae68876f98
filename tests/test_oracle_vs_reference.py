"""Pins the fp64 restatement (oracle/pas_oracle.c) against the UNMODIFIED reference CPU model
(oracle/_ref/libpas_ref.so = /root/reference/atmosphere/reference/functions.cc compiled in place +
oracle/ref_driver.cc) at the reference's own table sizes, on every pass of
atmosphere/reference/model.cc:140-237.

To stay within seconds, each 3-D pass is run by the reference on a strided subset of its texel
rows; the oracle is then given the reference's OWN (partially filled, otherwise zero) tables as
input and must reproduce the rows the reference computed to fp64 rounding. The full-size, fully
chained comparison (every texel of every table, 18 channels) is oracle/validate_oracle.py; its
result is recorded in tests/golden/earth18_meta.json ("oracle_vs_reference") and checked below.
CPU only; skipped where the prebuilt reference library is absent.
"""
import numpy as np
import pytest

from oracle import ref

pytestmark = pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libpas_ref.so not built")

MU, W = 128, 256


def rows_of(stride, total):
    return list(range(0, total, stride))


def max_rel(a, b):
    scale = np.abs(b).max()
    return float(np.abs(a - b).max() / scale) if scale > 0 else float(np.abs(a).max())


@pytest.fixture(scope="module")
def pair(pas, orc):
    spec = pas.earth(3, half_precision=True)
    cp = pas.channel_params(spec, [680.0, 550.0, 440.0])
    model = ref.RefModel(cp)
    o = orc.Oracle(cp)
    yield model, o
    model.close()


def test_every_pass_matches_the_reference(pair):
    model, o = pair
    # 2-D passes: every texel
    model.phase("transmittance")
    T_ref = model.read("transmittance")
    T = o.transmittance()
    assert max_rel(T, T_ref) < 1e-13
    model.phase("direct_irradiance")
    dE_ref = model.read("delta_irradiance")
    assert max_rel(o.direct_irradiance(T_ref), dE_ref) < 1e-13

    def check3(name, got, want, stride):
        g = got.reshape(got.shape[0], -1, W)[:, rows_of(stride, 32 * MU)]
        w = want.reshape(want.shape[0], -1, W)[:, rows_of(stride, 32 * MU)]
        assert np.abs(w).max() > 0, name
        assert max_rel(g, w) < 1e-12, name

    def oracle_rows(fn, stride, *args, **kw):
        """Runs an oracle pass on the same strided rows the reference computed."""
        out = None
        for row in rows_of(stride, 32 * MU):
            res = fn(*args, rows=(row, row + 1), **kw)
            if out is None:
                out = res
            elif isinstance(res, tuple):
                for a, b in zip(out, res):
                    a.reshape(a.shape[0], -1, W)[:, row] = b.reshape(b.shape[0], -1, W)[:, row]
            else:
                out.reshape(out.shape[0], -1, W)[:, row] = res.reshape(res.shape[0], -1, W)[:, row]
        return out

    s1, sd, sm = 8, 128, 32
    model.phase("single_scattering", stride=s1)
    dR_ref, dM_ref = model.read("delta_rayleigh"), model.read("delta_mie")
    dR, dM = oracle_rows(o.single_scattering, s1, T_ref)
    check3("delta_rayleigh", dR, dR_ref, s1)
    check3("delta_mie", dM, dM_ref, s1)

    zeros = np.zeros_like(dR_ref)
    model.phase("scattering_density", 2, stride=sd)
    dJ_ref = model.read("delta_density")
    dJ = oracle_rows(o.scattering_density, sd, T_ref, dR_ref, dM_ref, zeros, dE_ref, 2)
    check3("delta_density_2", dJ, dJ_ref, sd)

    # irradiance from order-1 radiance (reference/model.cc:204-215), every texel
    model.phase("indirect_irradiance", 2)
    dE2_ref = model.read("delta_irradiance")
    assert max_rel(o.indirect_irradiance(dR_ref, dM_ref, zeros, 1), dE2_ref) < 1e-12

    model.phase("multiple_scattering", 2, stride=sm)
    dS_ref = model.read("delta_multiple")
    dS = oracle_rows(lambda *a, **k: o.multiple_scattering(*a, **k)[0], sm, T_ref, dJ_ref)
    check3("delta_multiple_2", dS, dS_ref, sm)

    model.phase("scattering_density", 3, stride=sd)
    dJ3_ref = model.read("delta_density")
    dJ3 = oracle_rows(o.scattering_density, sd, T_ref, dR_ref, dM_ref, dS_ref, dE2_ref, 3)
    check3("delta_density_3", dJ3, dJ3_ref, sd)

    model.phase("indirect_irradiance", 3)
    dE3_ref = model.read("delta_irradiance")
    assert max_rel(o.indirect_irradiance(dR_ref, dM_ref, dS_ref, 2), dE3_ref) < 1e-12


def test_mappings_match_the_reference(pair):
    model, o = pair
    rng = np.random.default_rng(7)
    for _ in range(200):
        x, y, z = rng.uniform(0, 256), rng.uniform(0, 128), rng.uniform(0, 32)
        want = model.rmumusnu(x, y, z)
        got = o.rmumusnu_from_frag_coord(x, y, z)
        assert np.allclose(got, want, rtol=1e-12, atol=1e-12)
        r, mu, mu_s, nu, hit = want
        assert np.allclose(o.scattering_uvwz_from_rmumusnu(r, mu, mu_s, nu, hit),
                           model.uvwz(r, mu, mu_s, nu, hit), rtol=1e-10, atol=1e-12)


def test_recorded_full_size_validation(golden):
    """oracle/validate_oracle.py ran the restatement chained over 4 orders, 18 channels, every texel,
    against the reference run that produced tests/golden/: agreement to fp64 rounding."""
    _, _, meta = golden
    report = meta["oracle_vs_reference"]
    assert set(report) >= {"transmittance", "delta_rayleigh", "delta_density_4", "delta_multiple_4"}
    for name, m in report.items():
        assert m["max_abs_over_max"] < 1e-14, (name, m)
        assert m["max_rel"] < 1e-13, (name, m)

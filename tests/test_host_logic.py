"""Host-side logic that needs no GPU: the atmosphere definitions of the bench configs, the
wavelength grid and padding rules of atmosphere/model.cc, the r-slab partition, and the golden
error metric itself. CPU only."""
import math

import numpy as np
import pytest

from precomputed_atmospheric_scattering_b200 import atmospheres as atm
from precomputed_atmospheric_scattering_b200 import world
from tests import parity


def test_precomputed_wavelengths_follow_the_reference_grid():
    # model.cc:907-924
    assert atm.precomputed_wavelengths(3) == [680.0, 550.0, 440.0]
    lam = atm.precomputed_wavelengths(15)
    assert len(lam) == 15 and lam[0] == pytest.approx(360 + 0.5 * 470 / 15)
    assert np.allclose(np.diff(lam), 470 / 15)
    assert len(atm.precomputed_wavelengths(16)) == 18  # rounded up to a multiple of 3 (model.cc:917)


def test_interpolate_is_clamped_and_linear():
    # model.cc:535-552
    wl, v = [400.0, 500.0, 700.0], [1.0, 3.0, 7.0]
    assert atm.interpolate(wl, v, 300.0) == 1.0 and atm.interpolate(wl, v, 900.0) == 7.0
    assert atm.interpolate(wl, v, 450.0) == 2.0 and atm.interpolate(wl, v, 600.0) == 5.0


def test_earth_matches_the_demo_parameters():
    # demo.cc:188-284
    e = atm.earth(3, half_precision=True)
    assert len(e.wavelengths) == 48 and e.wavelengths[0] == 360.0 and e.wavelengths[-1] == 830.0
    assert e.bottom_radius == 6360000.0 and e.top_radius == 6420000.0
    assert e.max_sun_zenith_angle == pytest.approx(math.radians(102.0), rel=1e-6)
    assert atm.earth(3, half_precision=False).max_sun_zenith_angle == pytest.approx(math.radians(120.0), rel=1e-6)
    assert e.rayleigh_scattering[19] == pytest.approx(1.24062e-6 * 0.55 ** -4)
    assert e.mie_extinction[0] == pytest.approx(5.328e-3 / 1200.0)
    assert e.mie_scattering[0] == pytest.approx(0.9 * 5.328e-3 / 1200.0)
    assert e.absorption_extinction[24] == pytest.approx(300 * 2.687e20 / 15000 * 5.019e-25)
    assert len(e.absorption_density) == 2 and e.absorption_density[0].width == 25000.0


def test_channel_params_units_and_layer_padding():
    # model.cc:641-666, 718-734: lengths divided, inverse lengths multiplied by the length unit;
    # missing layers are zero layers inserted at the front
    e = atm.earth(3)
    cp = atm.channel_params(e, [680.0, 550.0, 440.0])
    assert cp.bottom_radius == 6360.0 and cp.top_radius == 6420.0
    assert cp.rayleigh_scattering[1] == pytest.approx(1.24062e-6 * 0.55 ** -4 * 1000.0)
    assert np.all(cp.profiles[0][0] == 0) and cp.profiles[0][1][2] == pytest.approx(-1000.0 / 8000.0)
    assert cp.profiles[2][0][0] == 25.0 and cp.profiles[2][0][3] == pytest.approx(1000.0 / 15000.0)
    assert cp.mu_s_min == pytest.approx(math.cos(e.max_sun_zenith_angle))
    with pytest.raises(ValueError):
        atm._pad_layers([atm.DensityProfileLayer()] * 3, 1000.0)


def test_r_slabs_cover_every_layer_once():
    for r_n in (32, 16, 7):
        for w in (1, 2, 3, 4, 8):
            ss = world.slabs(r_n, w)
            assert ss[0][0] == 0 and ss[-1][1] == r_n
            assert all(a[1] == b[0] for a, b in zip(ss, ss[1:]))
            assert max(e - b for b, e in ss) - min(e - b for b, e in ss) <= 1
    assert world.supported_world(32, 8) and not world.supported_world(32, 3)
    with pytest.raises(ValueError):
        world.slab(32, 2, 2)


def test_error_metric():
    ref = np.array([[1.0, 2.0, 0.0, 1e-9]])
    got = np.array([[1.001, 2.0, 1e-9, 2e-9]])
    m = parity.error_metrics(got, ref)
    assert m["max_rel"] == pytest.approx(1e-3, rel=1e-6)  # tiny references are masked out ...
    assert m["max_floor"] == pytest.approx(1e-3, rel=1e-6)  # ... but still bounded by the floor
    assert m["masked_fraction"] == 0.5 and m["nan"] == 0
    got[0, 2] = 1e-3
    assert parity.error_metrics(got, ref)["max_floor"] > 100


def test_luminance_matrices_match_the_reference_header(pas):
    """tests/golden/luminance.json was evaluated from the CIE table / XYZ_TO_SRGB matrix parsed out of
    the reference's atmosphere/constants.h (oracle/gen_luminance_golden.py); the product computes the
    same matrices from its own colorimetric data (csrc/cie1931.h). model.cc:907-943."""
    import json
    import os
    gold = json.load(open(os.path.join(parity.GOLDEN, "luminance.json")))
    for n in (15, 30):
        lam, L = pas.spectral_channels(n)
        assert np.allclose(lam, gold[f"n{n}"]["lambdas"], rtol=1e-15)
        assert np.array_equal(L, np.asarray(gold[f"n{n}"]["luminance_from_radiance"], dtype=np.float32))
    lam, L = pas.spectral_channels(3)
    assert list(lam) == [680.0, 550.0, 440.0] and np.array_equal(L, np.eye(3, dtype=np.float32))
    spec = pas.earth(3)
    rgb = pas.convert_spectrum_to_linear_srgb(spec.wavelengths, spec.solar_irradiance)
    assert np.allclose(rgb, gold["solar_srgb"], rtol=1e-12)


def test_ensemble_sweep_is_a_seeded_grid(pas):
    """BASELINE config 5: 64 atmospheres = turbidity x ozone x albedo; same seed, same atmospheres;
    the axes vary what the demo's keys vary (demo.cc:208-234) and nothing else."""
    a, b = pas.ensemble.sweep(), pas.ensemble.sweep()
    assert len(a) == 64
    key = lambda s: (s.mie_density[0].exp_scale, s.absorption_extinction[10], s.ground_albedo[0])
    assert [key(s) for s in a] == [key(s) for s in b]
    assert len({key(s) for s in a}) == 64
    assert [key(s) for s in pas.ensemble.sweep(seed=1)] != [key(s) for s in a]
    plain = pas.ensemble.sweep(2, 2, 2, jitter=0.0)
    assert len(plain) == 8
    heights = sorted({round(-1.0 / s.mie_density[0].exp_scale) for s in plain})
    assert heights == [800, 1200]
    base = pas.earth(3)
    for s in a:
        assert s.rayleigh_scattering == base.rayleigh_scattering and s.solar_irradiance == base.solar_irradiance
        assert 0.0 <= s.ground_albedo[0] <= 1.0
    with pytest.raises(ValueError):
        pas.ensemble.sweep(5, 1, 1)

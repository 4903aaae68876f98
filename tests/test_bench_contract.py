"""bench.py's host-side pieces, no GPU: the canonical work census of SURVEY.md section 8(d) (both the
builder and the judge use these numbers), the phase -> census key mapping, the reference-arm sampling
strides, and the ncu evidence file the roofline quotes."""
import json
import os

import pytest

import bench

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_canonical_work_matches_the_survey():
    # C = 15, K = 4: F ~ 6.6e11, U ~ 7.4e9; C = 3, K = 4: F ~ 2.0e11, U ~ 4.9e9; C = 3, K = 10: F ~ 5.3e11
    _, _, f, u = bench.canonical_work(15, 4)
    assert f == pytest.approx(6.6e11, rel=0.02) and u == pytest.approx(7.4e9, rel=0.02)
    _, _, f, u = bench.canonical_work(3, 4)
    assert f == pytest.approx(2.0e11, rel=0.03) and u == pytest.approx(4.9e9, rel=0.03)
    _, _, f, _ = bench.canonical_work(3, 10)
    assert f == pytest.approx(5.3e11, rel=0.03)
    flops, mufu, _, _ = bench.canonical_work(15, 4)
    # per-pass figures of DESIGN.md section 4
    assert flops["scattering_density_2"] == pytest.approx(226.6e9, rel=1e-3)
    assert flops["scattering_density_n"] == pytest.approx(148.2e9, rel=1e-3)
    assert flops["multiple_scattering"] == pytest.approx(37.4e9, rel=2e-3)
    assert flops["single_scattering"] == pytest.approx(22.2e9, rel=3e-3)
    assert set(mufu) == set(flops)


def test_phase_names_map_to_census_keys():
    flops, _, _, _ = bench.canonical_work()
    for name in ("transmittance", "single_scattering", "scattering_density_2", "scattering_density_3",
                 "scattering_density_10", "indirect_irradiance_2", "indirect_irradiance_4",
                 "multiple_scattering_2", "multiple_scattering_7"):
        assert bench.phase_key(name) in flops, name
    assert bench.phase_key("scattering_density_2") == "scattering_density_2"
    assert bench.phase_key("scattering_density_3") == "scattering_density_n"
    assert bench.phase_key("finalize") == "finalize"          # untimed bookkeeping phases pass through


def test_reference_arm_strides_visit_every_mu_row():
    # rows are k * 128 + j: an odd stride walks through every j
    for total, want in ((4096, 256), (4096, 64), (4096, 16)):
        s = bench._odd_stride(total, want)
        assert s % 2 == 1 and total // (want * 2) <= s <= total // want + 1
        assert len({r % 128 for r in range(0, total, s)}) >= min(128, total // s)


def test_ncu_evidence_file_names_the_kernels_the_bench_times():
    path = os.path.join(ROOT, "profiles", "ncu_hot_kernels.json")
    ncu = json.load(open(path))
    assert os.path.exists(os.path.join(ROOT, ncu["source"].split(" ")[0]))
    flops, _, _, _ = bench.canonical_work()
    for key in ("single_scattering", "scattering_density_2", "scattering_density_n", "multiple_scattering"):
        e = ncu["passes"][key]
        assert key in flops and e["traffic_bytes"] > 0 and 0 < e["fma_pipe_cycles_active_pct"] <= 100
        assert "kernel" in e and e["launches_averaged"] >= 1

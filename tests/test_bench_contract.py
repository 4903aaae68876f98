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


def test_ncu_census_gives_executed_fractions_below_one():
    """roofline.frac = executed fp32 flops (SASS opcode census of the ncu capture) / time / FMA peak:
    with the capture's own durations and the nominal 74.5 TFLOP/s no pass reaches the peak, and the
    ray marches report their real bound, the L1 data pipe (wavefronts per SM / elapsed cycles)."""
    ncu = json.load(open(os.path.join(ROOT, "profiles", "ncu_hot_kernels.json")))
    for key, e in ncu["passes"].items():
        ops = e["thread_instructions"]
        flop = (2 * ops.get("FFMA", 0) + 4 * ops.get("FFMA2", 0) + ops.get("FMUL", 0) + 2 * ops.get("FMUL2", 0) +
                ops.get("FADD", 0) + 2 * ops.get("FADD2", 0))
        assert flop == pytest.approx(e["fp32_flop_executed"], rel=1e-6)
        frac = e["fp32_flop_executed"] / (e["ms_under_ncu"] * 1e-3) / 74.5e12
        # (the ray setup pass is fp64 geometry, latency bound: a few fp32 operations only)
        assert (0.0 if key == "ray_setup" else 0.05) < frac < 1.0, (key, frac)
        # the executed fraction cannot exceed what the FMA pipe was busy
        assert frac <= e["fma_pipe_cycles_active_pct"] / 100 + 0.02, (key, frac)
        wf = e["l1_wavefronts_per_sm"] / e["sm_cycles_elapsed"]
        assert wf == pytest.approx(e["l1_data_pipe_wavefronts_pct"] / 100, abs=0.02), key
    assert "FFMA2" in ncu["passes"]["scattering_density_n"]["thread_instructions"]   # Blackwell packed fp32


def test_both_arms_print_the_same_config_object():
    for cfg in bench.CONFIGS:
        for n in (1, 2, 8):
            a, b = bench.config_dict(cfg, n), bench.config_dict(cfg, n)
            assert a == b and a["baseline_config"] == cfg and a["workload"] == bench.CONFIGS[cfg]["workload"]
            assert "l2" in a and ("1 GPU" if n == 1 else f"{n} GPUs") in a["parallelism"]
    assert bench.CONFIGS[2]["metric"] == "lut_precompute_ms_4_orders_15_wavelengths"
    # the driver's default line is config 2
    import argparse  # noqa: F401
    assert bench.CONFIGS[2]["wavelengths"] == 15 and bench.CONFIGS[2]["orders"] == 4


def test_parity_check_flags_a_wrong_table():
    """bench.py's `parity` key: the check passes on the reference's own numbers pushed through the
    product format and fails when one row of S is scaled by 1 %."""
    import numpy as np
    from tests import parity
    two, _, _ = parity.load_golden()
    rows = parity.load_rows()
    # RGB mode (config 1): a table whose rows reproduce the digests cannot be built without the full
    # reference tables, so exercise the failure path: zeros are out of tolerance everywhere
    S = np.zeros((32, 128, 256, 4), dtype=np.float32)
    E = np.moveaxis(two["irradiance"][15:18], 0, -1).astype(np.float32)
    T = np.moveaxis(two["transmittance"][15:18], 0, -1).astype(np.float32)
    m = bench.parity_check(1, S, np.concatenate([E, np.ones_like(E[..., :1])], -1),
                           np.concatenate([T, np.ones_like(T[..., :1])], -1), None, False)
    assert not m["ok"] and m["irradiance"] < 1e-6 and m["transmittance"] < 1e-6 and m["scattering_row_sums"] > 0.5
    assert m["n_texels"] == 32 * 128 * 256 * 4 + 3 * 16 * 64 + 3 * 64 * 256

"""Shared helpers of the parity tests: error metrics and golden access."""
from __future__ import annotations

import json
import os
from typing import Dict

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

# north_star tolerance: max relative error <= 1e-3 per texel, fp32 GPU vs fp64 CPU reference
REL_TOL = 1e-3
# texels whose reference magnitude is below REL_FLOOR * max|ref| (per table and channel) are
# compared with the absolute-floored metric instead (SURVEY.md section 7.2 "near-zero texels")
REL_FLOOR = 1e-6


def error_metrics(test: np.ndarray, ref: np.ndarray) -> Dict[str, float]:
    """test/ref shaped [C, ...]. Returns
    max_rel:   max |test/ref - 1| over texels with |ref| > REL_FLOOR * max_c|ref|
    max_floor: max |test - ref| / max(|ref|, REL_FLOOR * max_c|ref|) over ALL texels."""
    test = np.asarray(test, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert test.shape == ref.shape, (test.shape, ref.shape)
    C = ref.shape[0]
    t = test.reshape(C, -1)
    r = ref.reshape(C, -1)
    scale = np.abs(r).max(axis=1, keepdims=True)
    scale[scale == 0] = 1.0
    floor = REL_FLOOR * scale
    mask = np.abs(r) > floor
    rel = np.zeros_like(r)
    np.divide(np.abs(t - r), np.abs(r), out=rel, where=mask)
    floored = np.abs(t - r) / np.maximum(np.abs(r), floor)
    worst = np.unravel_index(np.argmax(floored), floored.shape)
    return {"max_rel": float(rel.max()), "max_floor": float(floored.max()),
            "worst_channel": int(worst[0]), "worst_texel": int(worst[1]),
            "masked_fraction": float(1.0 - mask.mean()), "nan": int(np.isnan(t).sum())}


def load_golden():
    two = np.load(os.path.join(GOLDEN, "earth18_2d.npz"))
    three = np.load(os.path.join(GOLDEN, "earth18_3d.npz"))
    meta = json.load(open(os.path.join(GOLDEN, "earth18_meta.json")))
    return two, three, meta


def sample3d(table: np.ndarray, indices: np.ndarray) -> np.ndarray:
    """table [C, R, MU, W] -> [C, nsamples] at golden indices (k, j, i)."""
    return table[:, indices[:, 0], indices[:, 1], indices[:, 2]]


# ---- per-row digests: every texel of every 3-D table (tests/golden/earth18_rows.npz) ------------------
# Written by oracle/gen_row_digest.py from the full-size run of the unmodified reference; one row = the
# 256 texels of one (layer k, mu row j), row index k * MU + j.

def row_weights(width: int) -> np.ndarray:
    """Position-dependent weights of the `wsum` digest (Knuth multiplicative hash of the column), in
    [0.5, 1.5): a permutation or a shift of the texels of a row changes the digest."""
    x = np.arange(width, dtype=np.uint64)
    h = (x * np.uint64(2654435761)) % np.uint64(1 << 32)
    return 0.5 + h.astype(np.float64) / float(1 << 32)


def row_digest(table: np.ndarray):
    """table [C, R, MU, W] -> (sum, wsum, max) over the texels of each row, each [C, R * MU] float64."""
    C, R, MU, W = table.shape
    t = np.asarray(table, dtype=np.float64).reshape(C, R * MU, W)
    return t.sum(axis=2), (t * row_weights(W)).sum(axis=2), t.max(axis=2)


def load_rows():
    return np.load(os.path.join(GOLDEN, "earth18_rows.npz"))


def digest_metrics(table: np.ndarray, rows, name: str, lanes=slice(None), floor: float = REL_FLOOR) -> Dict[str, float]:
    """Worst floored relative error of the three row digests of `table` [C, R, MU, W] against the
    reference digests `name` (channels `lanes`): |got - want| / max(|want|, floor * max_rows|want|),
    per channel and digest. Also the number of texels the digests cover."""
    got = row_digest(table)
    out = {"texels": int(np.prod(table.shape))}
    worst = 0.0
    for stat, g in zip(("sum", "wsum", "max"), got):
        want = np.asarray(rows[f"{name}/{stat}"][lanes], dtype=np.float64)
        assert want.shape == g.shape, (name, stat, want.shape, g.shape)
        scale = np.abs(want).max(axis=1, keepdims=True)
        scale[scale == 0] = 1.0
        err = np.abs(g - want) / np.maximum(np.abs(want), floor * scale)
        out[stat] = float(err.max())
        worst = max(worst, out[stat])
    out["worst"] = worst
    out["nan"] = int(np.isnan(np.asarray(table, dtype=np.float64)).sum())
    return out


# fp16 product tables: every texel is rounded to 11 significant bits (unit roundoff 2^-11) each time a
# pass accumulates into it: single scattering + one read-modify-write per multiple-scattering order = 4
# roundings at 4 orders. One texel (the row max) can be off by 4 * 2^-11 in the worst case; in the sums
# over the 256 texels of a row the roundings average out (3e-4 when the fp64 reference tables are pushed
# through the same fp16 accumulation).
HALF_TOL_TEXEL = 4.5 * 2.0 ** -11
HALF_TOL_SUM = 2.0 * 2.0 ** -11


def check_bench_product(S, E, T, L, two, rows, half_precision=True) -> Dict[str, float]:
    """The product tables of BASELINE config 2 (15 wavelengths, combined textures, 4 orders) against
    the reference run: S [R, MU, W, 4] against the per-row digests of the luminance table computed
    from the reference's fp64 tables (all R * MU * W texels, rgb + alpha), E [H, W, 4] against
    sum_n L . dE_n of the golden 2-D tables, T against the golden transmittance at 680/550/440 nm.
    Returns the worst floored relative errors; `ok` says whether they are within tolerance (fp16
    rounding for S when half_precision, the 1e-3 contract otherwise)."""
    S = np.asarray(S, dtype=np.float64)
    tol_sum, tol_max = (HALF_TOL_SUM, HALF_TOL_TEXEL) if half_precision else (REL_TOL, REL_TOL)
    m = digest_metrics(np.moveaxis(S, -1, 0), rows, "lum15_scattering", floor=1e-3)
    E_want = sum(np.tensordot(np.asarray(L, dtype=np.float64), two[f"delta_irradiance_{n}"][:15], axes=(1, 0))
                 for n in range(2, 5))
    e = error_metrics(np.moveaxis(np.asarray(E)[..., :3], -1, 0), E_want)
    t = error_metrics(np.moveaxis(np.asarray(T)[..., :3], -1, 0), two["transmittance"][15:18])
    s_sum = max(m["sum"], m["wsum"])
    out = {"scattering_row_sums": s_sum, "scattering_row_sums_tol": tol_sum,
           "scattering_row_max": m["max"], "scattering_row_max_tol": tol_max,
           "irradiance": e["max_floor"], "transmittance": t["max_floor"], "tol": REL_TOL,
           "n_texels": m["texels"] + int(E_want.size) + int(two["transmittance"][15:18].size),
           "nan": m["nan"] + e["nan"] + t["nan"]}
    # one headline figure on the scale of the 1e-3 contract: the worst error / its own tolerance
    out["max_floor"] = REL_TOL * max(out["irradiance"] / REL_TOL, out["transmittance"] / REL_TOL,
                                     s_sum / tol_sum, m["max"] / tol_max)
    out["ok"] = bool(out["nan"] == 0 and out["max_floor"] <= REL_TOL)
    return out

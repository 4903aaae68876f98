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

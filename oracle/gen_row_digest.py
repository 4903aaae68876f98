"""TEST INFRASTRUCTURE: per-row digest of EVERY texel of every 3-D table of the cached full-size run
of the UNMODIFIED reference CPU model (oracle/run_reference.py -> oracle/_cache/earth18/), so that
the parity tests pin all 1,048,576 texels of each table and not only the 3,240 sampled ones of
tests/golden/earth18_3d.npz.

  tests/golden/earth18_rows.npz
    <table>/sum, <table>/wsum, <table>/max   float32 [18 channels][32 * 128 rows]
        for the 9 per-channel intermediates (delta_rayleigh, delta_mie, delta_density_n,
        delta_multiple_n) and the per-channel sum `scattering`; one row = the 256 texels
        (nu slab, mu_s) of one (layer k, mu row j); row index = k * 128 + j.
        sum  = sum_x v[x]              (computed in float64)
        wsum = sum_x w[x] v[x]         w[x] = tests/parity.py:row_weights(256): a fixed pseudo-random weight in
                                       [0.5, 1.5) per column, so that a permutation or a shift of
                                       the texels of a row changes the digest
        max  = max_x v[x]
    lum15_scattering/{sum,wsum,max}          float32 [4][4096]: the combined luminance product of
        BASELINE config 2 (15 wavelengths): rgb = sum_c L[a][c] (dR + sum_n dS_n / P_R(nu))[c],
        alpha = (L . dM)_red (atmosphere/model.cc:142-157, 192-208), from the fp64 tables and the
        committed luminance matrix (tests/golden/luminance.json).

Channels 0..14 are the 15 spectral channels of BASELINE config 2, channels 15..17 the RGB channels
(680/550/440 nm) of config 1. Usage: python oracle/gen_row_digest.py"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
CACHE = os.path.join(HERE, "_cache", "earth18")
OUT = os.path.join(ROOT, "tests", "golden")


from tests.parity import row_digest as digest  # noqa: E402  (the checker uses the same function)


def main():
    from oracle import oracle as orc
    from oracle.run_reference import earth18_channels

    meta = json.load(open(os.path.join(CACHE, "meta.json")))
    orders = meta["orders"]
    names = ["delta_rayleigh", "delta_mie", "scattering"]
    for n in range(2, orders + 1):
        names += [f"delta_density_{n}", f"delta_multiple_{n}"]
    out = {}
    for name in names:
        a = np.load(os.path.join(CACHE, name + ".npy"), mmap_mode="r")
        s, ws, mx = digest(a)
        out[name + "/sum"], out[name + "/wsum"], out[name + "/max"] = (
            s.astype(np.float32), ws.astype(np.float32), mx.astype(np.float32))
        print(name, a.shape, flush=True)
    # the combined luminance product of config 2 from the fp64 tables
    L = np.asarray(json.load(open(os.path.join(OUT, "luminance.json")))["n15"]["luminance_from_radiance"])
    _, cp = earth18_channels()
    orc.build()
    nu = orc.Oracle(cp).texel_params()[..., 3]
    inter = {"nu": nu}
    for name in ("delta_rayleigh", "delta_mie"):
        inter[name] = np.load(os.path.join(CACHE, name + ".npy"), mmap_mode="r")[:15]
    S = np.tensordot(L, inter["delta_rayleigh"], axes=(1, 0))
    for n in range(2, orders + 1):
        dS = np.load(os.path.join(CACHE, f"delta_multiple_{n}.npy"), mmap_mode="r")[:15]
        S = S + np.tensordot(L, dS, axes=(1, 0)) / orc.rayleigh_phase(nu)[None]
    A = np.tensordot(L[0], inter["delta_mie"], axes=(0, 0))
    s, ws, mx = digest(np.concatenate([S, A[None]], axis=0))
    out["lum15_scattering/sum"], out["lum15_scattering/wsum"], out["lum15_scattering/max"] = (
        s.astype(np.float32), ws.astype(np.float32), mx.astype(np.float32))
    path = os.path.join(OUT, "earth18_rows.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path))


if __name__ == "__main__":
    main()

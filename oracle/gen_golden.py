"""TEST INFRASTRUCTURE: turns the cached full-size run of the UNMODIFIED reference CPU model
(oracle/run_reference.py -> oracle/_cache/earth18/) into the small committed fixtures under
tests/golden/:

  earth18_2d.npz       every texel of the 2-D tables (transmittance, delta_irradiance_1..N,
                       irradiance) for the 18 channels, float64.
  earth18_3d.npz       a fixed sample of 3-D texels (seeded random + all table corners / halves /
                       first & last layers) of every 3-D table (delta_rayleigh, delta_mie,
                       delta_density_n, delta_multiple_n, scattering), float64, with their indices.
  earth18_meta.json    wavelengths, sizes, reference phase timings, oracle-vs-reference report.

Channels 0..14 are the 15 spectral channels of BASELINE config 2, channels 15..17 the RGB channels
(680/550/440 nm) of config 1. Usage: python oracle/gen_golden.py"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CACHE = os.path.join(HERE, "_cache", "earth18")
OUT = os.path.join(ROOT, "tests", "golden")
N_RANDOM = 3000
SEED = 20261017


def sample_indices(shape):
    R, MU, W = shape
    rng = np.random.default_rng(SEED)
    idx = set()
    ks = [0, 1, R // 2, R - 2, R - 1]
    js = [0, 1, MU // 2 - 1, MU // 2, MU // 2 + 1, MU - 1]
    is_ = [0, 1, 31, 32, W // 2, W - 33, W - 32, W - 1]
    for k in ks:
        for j in js:
            for i in is_:
                idx.add((k, j, i))
    flat = rng.choice(R * MU * W, size=N_RANDOM, replace=False)
    for f in flat:
        idx.add((int(f // (MU * W)), int((f // W) % MU), int(f % W)))
    arr = np.array(sorted(idx), dtype=np.int32)
    return arr


def main():
    os.makedirs(OUT, exist_ok=True)
    meta = json.load(open(os.path.join(CACHE, "meta.json")))
    rep = os.path.join(CACHE, "oracle_vs_reference.json")
    if os.path.exists(rep):
        meta["oracle_vs_reference"] = json.load(open(rep))
    orders = meta["orders"]
    two_d = {"transmittance": np.load(os.path.join(CACHE, "transmittance.npy")),
             "irradiance": np.load(os.path.join(CACHE, "irradiance.npy"))}
    for n in range(1, orders + 1):
        two_d[f"delta_irradiance_{n}"] = np.load(os.path.join(CACHE, f"delta_irradiance_{n}.npy"))
    np.savez_compressed(os.path.join(OUT, "earth18_2d.npz"), **two_d)
    names = ["delta_rayleigh", "delta_mie", "scattering"]
    for n in range(2, orders + 1):
        names += [f"delta_density_{n}", f"delta_multiple_{n}"]
    first = np.load(os.path.join(CACHE, names[0] + ".npy"), mmap_mode="r")
    idx = sample_indices(first.shape[1:])
    three_d = {"indices": idx}
    for name in names:
        a = np.load(os.path.join(CACHE, name + ".npy"), mmap_mode="r")
        three_d[name] = np.ascontiguousarray(a[:, idx[:, 0], idx[:, 1], idx[:, 2]])
    np.savez_compressed(os.path.join(OUT, "earth18_3d.npz"), **three_d)
    meta["golden_3d_samples"] = int(len(idx))
    meta["generator"] = "oracle/run_reference.py + oracle/gen_golden.py (unmodified reference CPU model)"
    json.dump(meta, open(os.path.join(OUT, "earth18_meta.json"), "w"), indent=1)
    for f in os.listdir(OUT):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()

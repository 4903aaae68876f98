"""TEST INFRASTRUCTURE: runs the unmodified reference CPU model (oracle/_ref) on the Earth/demo
atmosphere with 18 lanes = the 15 spectral channels of BASELINE config 2 followed by the RGB
channels (680/550/440 nm) of config 1, 4 scattering orders, and caches every table as float64
under oracle/_cache/earth18/ (git- and gpurun-ignored; ~2.5 GB). oracle/gen_golden.py turns the
cache into the small committed fixtures under tests/golden/.

Usage: python oracle/run_reference.py [--orders 4] [--out oracle/_cache/earth18]"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref  # noqa: E402
from precomputed_atmospheric_scattering_b200 import atmospheres as atm  # noqa: E402


def earth18_channels():
    spec = atm.earth(15, half_precision=True)  # demo default: half precision => 102 deg
    lambdas = atm.precomputed_wavelengths(15) + [atm.LAMBDA_R, atm.LAMBDA_G, atm.LAMBDA_B]
    return spec, atm.channel_params(spec, lambdas)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--orders", type=int, default=4)
    ap.add_argument("--out", default=os.path.join(os.path.dirname(__file__), "_cache", "earth18"))
    args = ap.parse_args()
    os.makedirs(args.out, exist_ok=True)
    spec, cp = earth18_channels()
    model = ref.RefModel(cp)
    t0 = time.time()

    def dump(name, arr):
        np.save(os.path.join(args.out, name + ".npy"), arr)

    times = model.precompute(args.orders, dump=dump, log=lambda s: print(s, flush=True))
    meta = {"lambdas": list(map(float, cp.lambdas)), "orders": args.orders, "phase_seconds": times,
            "total_seconds": time.time() - t0, "threads": model.nthreads, "sizes": model.sz,
            "lanes_computed": 47}
    with open(os.path.join(args.out, "meta.json"), "w") as f:
        json.dump(meta, f, indent=1)
    print(json.dumps(meta["phase_seconds"]))


if __name__ == "__main__":
    main()

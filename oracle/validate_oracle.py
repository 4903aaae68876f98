"""TEST INFRASTRUCTURE: pins the fp64 restatement (oracle/pas_oracle.c) against the unmodified
reference CPU model at FULL size: runs the restatement chained over all orders on the same 18
channels as oracle/run_reference.py and reports, per table, the max relative difference with the
cached reference tables (oracle/_cache/earth18). Result is recorded in DESIGN.md.

Usage: python oracle/validate_oracle.py [--orders 4]"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle  # noqa: E402
from oracle.run_reference import earth18_channels  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--orders", type=int, default=4)
    ap.add_argument("--cache", default=os.path.join(os.path.dirname(__file__), "_cache", "earth18"))
    args = ap.parse_args()
    _, cp = earth18_channels()
    o = oracle.Oracle(cp)
    report = {}

    def check(name, arr):
        ref = np.load(os.path.join(args.cache, name + ".npy"), mmap_mode="r")
        ref = np.asarray(ref)
        scale = np.abs(ref).max(axis=tuple(range(1, ref.ndim)), keepdims=True)
        m = np.abs(ref) > 1e-30
        rel = np.zeros_like(ref)
        rel[m] = np.abs(arr[m] / ref[m] - 1.0)
        report[name] = {"max_rel": float(rel.max()),
                        "max_abs_over_max": float((np.abs(arr - ref) / scale).max()),
                        "ref_zeros": int((~m).sum()),
                        "oracle_nonzero_where_ref_zero": int((arr[~m] != 0).sum())}
        print(name, report[name], flush=True)

    o.precompute(args.orders, dump=check, log=lambda s: print(s, flush=True))
    with open(os.path.join(args.cache, "oracle_vs_reference.json"), "w") as f:
        json.dump(report, f, indent=1)


if __name__ == "__main__":
    main()

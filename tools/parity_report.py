"""GPU parity report against the committed golden fixtures (tests/golden, produced by the
unmodified reference CPU model): runs the CUDA path on the Earth atmosphere in spectral (15
wavelengths) and RGB mode with intermediate capture, and prints per-table error metrics and the
per-phase device timings. Needs a GPU. Usage: python tools/parity_report.py [--orders 4]"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import precomputed_atmospheric_scattering_b200 as pas  # noqa: E402
from tests import parity  # noqa: E402


def report(model, lanes, two, three, orders, label, out):
    idx = three["indices"]
    names2 = ["transmittance"] + [f"delta_irradiance_{n}" for n in range(1, orders + 1)]
    names3 = ["delta_rayleigh", "delta_mie"]
    for n in range(2, orders + 1):
        names3 += [f"delta_density_{n}", f"delta_multiple_{n}"]
    for name in names2:
        got = model.intermediate(name)
        m = parity.error_metrics(got, two[name][lanes])
        out[f"{label}/{name}"] = m
        print(f"{label:8s} {name:22s} max_rel={m['max_rel']:.3e} max_floor={m['max_floor']:.3e} "
              f"worst=(c{m['worst_channel']}, t{m['worst_texel']}) nan={m['nan']}", flush=True)
    for name in names3:
        got = parity.sample3d(model.intermediate(name), idx)
        m = parity.error_metrics(got, three[name][lanes])
        k, j, i = idx[m["worst_texel"]]
        out[f"{label}/{name}"] = m
        print(f"{label:8s} {name:22s} max_rel={m['max_rel']:.3e} max_floor={m['max_floor']:.3e} "
              f"worst=(c{m['worst_channel']}, k{k} j{j} i{i}) nan={m['nan']}", flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--orders", type=int, default=4)
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    two, three, meta = parity.load_golden()
    out = {}
    for label, n, lanes in (("rgb", 3, slice(15, 18)), ("spectral", 15, slice(0, 15))):
        spec = pas.earth(n, half_precision=False, max_sun_zenith_deg=102.0)
        model = pas.Model.from_spec(spec)
        model.set_capture(True)
        t0 = time.time()
        model.Init(args.orders)
        print(f"{label}: Init wall {1e3 * (time.time() - t0):.2f} ms (with capture copies)")
        report(model, lanes, two, three, args.orders, label, out)
        if n == 3:
            # final radiance tables in RGB mode are directly comparable (L = identity)
            S = model.scattering
            got = np.moveaxis(S[..., :3], -1, 0)
            m = parity.error_metrics(parity.sample3d(got, three["indices"]), three["scattering"][lanes])
            out["rgb/scattering"] = m
            print(f"rgb      scattering             max_rel={m['max_rel']:.3e} max_floor={m['max_floor']:.3e}")
            E = np.moveaxis(model.irradiance[..., :3], -1, 0)
            m = parity.error_metrics(E, two["irradiance"][lanes])
            out["rgb/irradiance"] = m
            print(f"rgb      irradiance             max_rel={m['max_rel']:.3e} max_floor={m['max_floor']:.3e}")
        model.set_capture(False)
        for _ in range(3):
            model.Init(args.orders)
        t0 = time.time()
        model.Init(args.orders)
        wall = 1e3 * (time.time() - t0)
        tm = model.last_timings()
        print(f"{label}: Init wall {wall:.3f} ms, device phases sum {sum(tm.values()):.3f} ms, "
              f"launches {model.last_launch_count()}")
        print(json.dumps({k: round(v, 4) for k, v in tm.items()}))
        out[f"{label}/timings_ms"] = tm
        out[f"{label}/wall_ms"] = wall
        model.close()
    if args.json:
        json.dump(out, open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()

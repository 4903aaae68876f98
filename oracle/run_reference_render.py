"""TEST INFRASTRUCTURE: golden IMAGES of the reference's integration test.

Runs the UNMODIFIED reference CPU model (oracle/_ref) exactly like
atmosphere/reference/model_test.cc does for its CPU images: the test atmosphere
(model_test.cc:222-308), all 47 spectral lanes at 360 + 10 i nm, 4 scattering orders, then the test
scene of model_test.glsl rendered by the reference's own GetViewRayRadiance for the two sun
positions of the test cases (zenith 65 and 88 degrees, azimuth 90). Per pixel the 47-lane radiance is
reduced the way RenderCpuImage does (model_test.cc:669-738):
  radiance   = lanes 32 / 19 / 8 (680 / 550 / 440 nm), grass / snow scene albedos
  luminance  = MAX_LUMINOUS_EFFICACY * XYZ_TO_SRGB * sum_lanes(radiance * cie_xyz_bar) * 10 nm,
               once with the grass / snow albedos and once with the constant 0.1 / 0.8 albedos of
               the *ConstantAlbedo test cases (model_test.cc:873-874)
and both are stored BEFORE tone mapping, float32, in tests/golden/render_earth47.npz, at the
reference's 640x360 frame subsampled 2x (pixel centres of a 320x180 frame use the same view matrix
convention, model_from_clip built for the image size actually rendered).

The full tables are cached under oracle/_cache/earth47/ (git- and gpurun-ignored, ~5 GB as float64)
so that the images can be regenerated at another size without repeating the 8-minute precompute.

Usage: python oracle/run_reference_render.py [--width 320 --height 180] [--reuse]"""
import argparse
import json
import os
import re
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402
from precomputed_atmospheric_scattering_b200 import atmospheres as atm  # noqa: E402
from precomputed_atmospheric_scattering_b200 import scene as scn  # noqa: E402

REFERENCE = os.environ.get("PAS_REFERENCE", "/root/reference")
LANES = 47
XYZ_TO_SRGB = np.array([[3.2406, -1.5372, -0.4986], [-0.9689, 1.8758, 0.0415], [0.0557, -0.2040, 1.0570]])
MAX_LUMINOUS_EFFICACY = 683.0
SUNS = [(65.0, 90.0), (88.0, 90.0)]


def lane_wavelengths():
    return [360.0 + 10.0 * i for i in range(LANES)]


def cie_table():
    """CIE_2_DEG_COLOR_MATCHING_FUNCTIONS parsed from the reference header (atmosphere/constants.h)."""
    text = open(os.path.join(REFERENCE, "atmosphere", "constants.h")).read()
    body = text[text.index("CIE_2_DEG_COLOR_MATCHING_FUNCTIONS[380]"):]
    body = body[body.index("{") + 1:body.index("};")]
    vals = [float(v) for v in re.findall(r"[-+]?\d+\.?\d*(?:[eE][-+]?\d+)?", body)]
    assert len(vals) == 380
    return np.array(vals).reshape(95, 4)


def cie_on_lanes():
    """x/y/z bar resampled onto the 47 lanes like DimensionlessSpectrum(wavelengths, values)
    (model_test.cc:672-683; scalar_function.h:225-259: linear, constant outside)."""
    t = cie_table()
    lam = np.array(lane_wavelengths())
    return np.stack([np.interp(lam, t[:, 0], t[:, k]) for k in (1, 2, 3)])  # [3, 47]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--width", type=int, default=320)
    ap.add_argument("--height", type=int, default=180)
    ap.add_argument("--orders", type=int, default=4)
    ap.add_argument("--reuse", action="store_true", help="reuse oracle/_cache/earth47 tables")
    ap.add_argument("--cache", default=os.path.join(HERE, "_cache", "earth47"))
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden", "render_earth47.npz"))
    args = ap.parse_args()
    os.makedirs(args.cache, exist_ok=True)

    spec = atm.model_test_earth()
    cp = atm.channel_params(spec, lane_wavelengths())
    model = ref.RefModel(cp)
    names = {"transmittance": "transmittance", "scattering": "scattering", "delta_mie": "delta_mie",
             "irradiance": "irradiance"}
    t0 = time.time()
    if args.reuse and all(os.path.exists(os.path.join(args.cache, n + ".npy")) for n in names):
        for n in names:
            model.write(n, np.load(os.path.join(args.cache, n + ".npy")))
        times = json.load(open(os.path.join(args.cache, "meta.json")))["phase_seconds"]
    else:
        times = model.precompute(args.orders, log=lambda s: print(s, flush=True))
        for n in names:
            np.save(os.path.join(args.cache, n + ".npy"), model.read(n))
        json.dump({"phase_seconds": times, "orders": args.orders, "threads": model.nthreads},
                  open(os.path.join(args.cache, "meta.json"), "w"), indent=1)
    print(f"tables ready in {time.time() - t0:.1f} s", flush=True)

    lam = lane_wavelengths()
    cie = cie_on_lanes()
    out = {"lane_wavelengths": np.array(lam), "suns": np.array(SUNS),
           "size": np.array([args.width, args.height])}
    # scene albedos: the fixture's grass / snow spectra (model_test.cc:317-318), or the constants
    # 0.1 / 0.8 of the *ConstantAlbedo test cases (model_test.cc:873-874)
    albedos = {"spectral": (np.array([scn.grass_albedo(l) for l in lam]),
                            np.array([scn.snow_albedo(l) for l in lam])),
               "constant": (np.full(LANES, 0.1), np.full(LANES, 0.8))}
    for zen, az in SUNS:
        view = scn.model_test_view(zen, az, False, width=args.width, height=args.height,
                                   sun_angular_radius=spec.sun_angular_radius)
        tag = f"sun{int(zen)}"
        for kind, (ground, sphere) in albedos.items():
            img = model.render_scene(view, ground, sphere)  # [H, W, 47]
            print(f"sun {zen} {kind} albedo: rendered in {model.last_render_seconds:.1f} s", flush=True)
            if kind == "spectral":
                out[f"{tag}_radiance"] = img[..., [32, 19, 8]].astype(np.float32)
            xyz = np.einsum("hwl,kl->hwk", img, cie) * 10.0 * MAX_LUMINOUS_EFFICACY
            out[f"{tag}_luminance_{kind}"] = np.einsum("hwk,ck->hwc", xyz, XYZ_TO_SRGB).astype(np.float32)
    np.savez_compressed(args.out, **out)
    meta = {"generator": "oracle/run_reference_render.py (unmodified reference CPU model + model_test.glsl)",
            "phase_seconds": times, "width": args.width, "height": args.height, "orders": args.orders,
            "atmosphere": "reference/model_test.cc:222-308 (sun radius 0.2678 deg, mu_s_min cos 102 deg, albedo 0.1)",
            "scene": "grass ground, snow sphere (model_test.cc:317-355), camera model_test.cc:436-477"}
    json.dump(meta, open(args.out.replace(".npz", "_meta.json"), "w"), indent=1)
    print(args.out, os.path.getsize(args.out))


if __name__ == "__main__":
    main()

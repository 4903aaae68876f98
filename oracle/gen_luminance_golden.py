"""TEST INFRASTRUCTURE: writes tests/golden/luminance.json -- the wavelength grids and
luminance_from_radiance matrices of atmosphere/model.cc:907-943 for n = 15 and n = 30 precomputed
wavelengths, and the SKY/SUN_SPECTRAL_RADIANCE_TO_LUMINANCE factors of model.cc:562-595 for the
Earth/demo solar spectrum -- evaluated in numpy from the CIE table and XYZ_TO_SRGB matrix PARSED OUT
OF the reference header /root/reference/atmosphere/constants.h at generation time (nothing is
copied into the repo but the resulting numbers). Usage: python oracle/gen_luminance_golden.py"""
import json
import math
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from precomputed_atmospheric_scattering_b200 import atmospheres as atm  # noqa: E402

REF = os.environ.get("PAS_REFERENCE", "/root/reference")


def parse_array(text, name):
    body = re.search(name + r"\[\d+\]\s*=\s*\{(.*?)\};", text, flags=re.S).group(1)
    body = re.sub(r"//.*", "", body)
    return np.array([float(v) for v in re.findall(r"[-+]?\d*\.?\d+(?:[eE][-+]?\d+)?", body)])


def main():
    text = open(os.path.join(REF, "atmosphere", "constants.h")).read()
    cie = parse_array(text, "CIE_2_DEG_COLOR_MATCHING_FUNCTIONS").reshape(-1, 4)
    m = parse_array(text, "XYZ_TO_SRGB").reshape(3, 3)
    assert cie.shape == (95, 4) and cie[0, 0] == 360 and cie[-1, 0] == 830

    def cie_value(lam, col):  # model.cc:521-533
        if lam <= 360 or lam >= 830:
            return 0.0
        u = (lam - 360) / 5.0
        row = int(math.floor(u))
        u -= row
        return cie[row, col] * (1 - u) + cie[row + 1, col] * u

    out = {"generator": "oracle/gen_luminance_golden.py from atmosphere/constants.h:69-173"}
    for n in (15, 30):
        lam = atm.precomputed_wavelengths(n)
        dl = 470.0 / len(lam)
        L = np.zeros((3, len(lam)))
        for j, l in enumerate(lam):
            xyz = np.array([cie_value(l, c) for c in (1, 2, 3)])
            L[:, j] = np.float32((m @ xyz) * dl)  # float cast, model.cc:934-942
        out[f"n{n}"] = {"lambdas": lam, "luminance_from_radiance": L.tolist()}
    spec = atm.earth(3)
    for key, power in (("sky_k", -3.0), ("sun_k", 0.0)):  # model.cc:562-595
        k = np.zeros(3)
        lam_rgb = [680.0, 550.0, 440.0]
        sol_rgb = [atm.interpolate(spec.wavelengths, spec.solar_irradiance, l) for l in lam_rgb]
        for lam in range(360, 830):
            xyz = np.array([cie_value(lam, c) for c in (1, 2, 3)])
            bar = m @ xyz
            irr = atm.interpolate(spec.wavelengths, spec.solar_irradiance, lam)
            for a in range(3):
                k[a] += bar[a] * irr / sol_rgb[a] * (lam / lam_rgb[a]) ** power
        out[key] = (683.0 * k).tolist()
    # ConvertSpectrumToLinearSrgb (model.cc:1020-1040) of the Earth solar spectrum
    xyz = np.zeros(3)
    for lam in range(360, 830):
        v = atm.interpolate(spec.wavelengths, spec.solar_irradiance, lam)
        xyz += np.array([cie_value(lam, c) for c in (1, 2, 3)]) * v
    out["solar_srgb"] = (683.0 * (m @ xyz)).tolist()
    path = os.path.join(ROOT, "tests", "golden", "luminance.json")
    json.dump(out, open(path, "w"), indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()

"""The C++ drop-in atmosphere::Model (include/atmosphere_b200/model.h): compiles against the C ABI
with the reference's constructor / Init / SetProgramUniforms signatures (CPU test), and produces
byte-identical tables to the Python mirror when run on a GPU (gpu test)."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "model_shim_main.cc")


def build(tmp_path, pas):
    exe = str(tmp_path / "model_shim_main")
    pkg = os.path.dirname(pas.LIB_PATH)
    subprocess.check_call(["g++", "-std=c++14", "-Wall", "-Werror", "-O1",
                           "-I" + os.path.join(ROOT, "include", "atmosphere_b200"), SRC, "-o", exe,
                           "-L" + pkg, "-l:libpas_b200.so", "-Wl,-rpath," + pkg])
    return exe


def write_spectra(path, spec):
    with open(path, "w") as f:
        for row in zip(spec.wavelengths, spec.solar_irradiance, spec.rayleigh_scattering,
                       spec.mie_scattering, spec.mie_extinction, spec.absorption_extinction,
                       spec.ground_albedo):
            f.write(" ".join(repr(float(v)) for v in row) + "\n")


def test_shim_compiles_and_fails_loudly_without_gpu(tmp_path, pas):
    exe = build(tmp_path, pas)
    import torch
    if torch.cuda.is_available():
        pytest.skip("the run is covered by the gpu test")
    write_spectra(tmp_path / "spectra.txt", pas.earth(3, half_precision=True))
    p = subprocess.run([exe, str(tmp_path), str(tmp_path / "spectra.txt"), "3"], capture_output=True, text=True)
    assert p.returncode == 1 and "CUDA" in p.stderr   # no CPU fallback


@pytest.mark.gpu
@pytest.mark.parametrize("n", [3, 15])
def test_shim_tables_equal_python_mirror(tmp_path, pas, n):
    exe = build(tmp_path, pas)
    spec = pas.earth(n, half_precision=True)
    write_spectra(tmp_path / "spectra.txt", spec)
    out = subprocess.run([exe, str(tmp_path), str(tmp_path / "spectra.txt"), str(n)], capture_output=True,
                         text=True, check=True).stdout
    assert "scattering 256x128x32 bytes_per_channel 2" in out and "shader 0" in out
    model = pas.Model.from_spec(spec)
    model.Init(4)
    for fn, tab in (("transmittance.dat", model.transmittance), ("scattering.dat", model.scattering),
                    ("irradiance.dat", model.irradiance)):
        raw = np.fromfile(os.path.join(tmp_path, fn), dtype="<f4")
        assert np.array_equal(raw.reshape(tab.shape), tab), fn
    model.close()

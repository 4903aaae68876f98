"""The C-ABI shared library: loads without a GPU, exports exactly what include/pas_b200.h declares,
validates arguments like the reference documents them, and fails loudly (no CPU fallback) when no
CUDA device is present. CPU only -- no compute calls."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "pas_b200.h")


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pas_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(pas):
    lib = pas.load_library()
    names = declared_functions()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), f"{name} is declared in include/pas_b200.h but not exported"
    assert lib.pas_abi_version() == 1


def test_every_entry_point_cites_the_reference():
    text = open(HEADER).read()
    for needle in ("atmosphere/model.cc:613-795", "atmosphere/model.cc:866-975", "atmosphere/model.h:182-281",
                   "atmosphere/model.cc:1020-1040", "atmosphere/demo/webgl/precompute.cc"):
        assert needle in text


def test_library_does_not_link_the_oracle(pas):
    """The product path must not route through the CPU checker."""
    blob = open(pas.LIB_PATH, "rb").read()
    assert b"paso_" not in blob and b"liboracle" not in blob and b"libpas_ref" not in blob
    import precomputed_atmospheric_scattering_b200 as pkg
    pkg_dir = os.path.dirname(pkg.__file__)
    for fn in os.listdir(pkg_dir):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg_dir, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), fn


def _no_gpu():
    import torch
    return not torch.cuda.is_available()


def test_argument_validation(pas):
    spec = pas.earth(3)
    with pytest.raises(ValueError):  # model.cc:539: one value per wavelength
        pas.Model(spec.wavelengths, spec.solar_irradiance[:-1], *[None] * 17)
    bad = pas.earth(3)
    bad.top_radius = bad.bottom_radius
    with pytest.raises(pas.PasError) as e:
        pas.Model.from_spec(bad)
    assert e.value.status == 1 and "bottom_radius" in str(e.value)
    bad = pas.earth(3)
    bad.sun_angular_radius = 0.2  # model.h:194-195
    with pytest.raises(pas.PasError) as e:
        pas.Model.from_spec(bad)
    assert e.value.status == 1
    bad = pas.earth(3)
    bad.wavelengths = list(reversed(bad.wavelengths))
    with pytest.raises(pas.PasError) as e:
        pas.Model.from_spec(bad)
    assert e.value.status == 1 and "increasing" in str(e.value)
    bad = pas.earth(3)
    bad.rayleigh_density = bad.rayleigh_density * 3  # model.h:206-207: at most 2 layers
    with pytest.raises(pas.PasError) as e:
        pas.Model.from_spec(bad)
    assert e.value.status == 1


@pytest.mark.skipif(not _no_gpu(), reason="only meaningful on a box without a GPU")
def test_fails_loudly_without_a_gpu(pas):
    with pytest.raises(pas.PasError) as e:
        pas.Model.from_spec(pas.earth(3))
    assert e.value.status == 2  # PAS_ERR_CUDA: there is no CPU fallback
    assert "CUDA" in str(e.value)


def test_convert_spectrum_to_linear_srgb(pas):
    """atmosphere/model.cc:1020-1040: sum over 360..829 nm of cie(lambda) * spectrum, times
    XYZ->sRGB and 683 lm/W. A flat spectrum is integrated against the colour-matching functions:
    Y integrates to ~106.86 (sum of y_bar at 1 nm steps)."""
    wl = [360.0, 830.0]
    r, g, b = pas.convert_spectrum_to_linear_srgb(wl, [1.0, 1.0])
    assert r > 0 and g > 0 and b > 0
    # linearity
    r2, g2, b2 = pas.convert_spectrum_to_linear_srgb(wl, [2.0, 2.0])
    assert (r2, g2, b2) == pytest.approx((2 * r, 2 * g, 2 * b), rel=1e-12)
    # luminance Y = 0.2126 R + 0.7152 G + 0.0722 B of a flat unit spectrum = 683 * sum(y_bar)
    y = 0.2126 * r + 0.7152 * g + 0.0722 * b
    assert y == pytest.approx(683.0 * 106.86, rel=2e-3)


def test_last_error_is_thread_local_string(pas):
    lib = pas.load_library()
    assert lib.pas_model_init(None, 4) == 1
    assert b"NULL" in lib.pas_last_error()


def test_null_handles_are_rejected_everywhere(pas):
    """Every entry point that takes a model validates it before touching CUDA: status 1
    (PAS_ERR_INVALID_ARGUMENT) and a message, never a crash -- also without a GPU."""
    import ctypes
    lib = pas.load_library()
    n = ctypes.c_size_t(0)
    buf = ctypes.create_string_buffer(pas.model.IPC_EXPORT_BYTES)
    calls = [
        lambda: lib.pas_model_init(None, 4),
        lambda: lib.pas_model_init_async(None, 4),
        lambda: lib.pas_model_wait(None),
        lambda: lib.pas_model_ipc_export(None, 0, 2, buf, ctypes.byref(n)),
        lambda: lib.pas_model_attach_peers(None, buf, pas.model.IPC_EXPORT_BYTES),
        lambda: lib.pas_model_attach_world(None, 0, 2, None),
        lambda: lib.pas_model_save_dat(None, b"/tmp"),
        lambda: lib.pas_model_set_capture(None, 1),
    ]
    for call in calls:
        assert call() == 1
        assert len(lib.pas_last_error()) > 0
    lib.pas_model_destroy(None)  # a no-op, like free(NULL)


def test_ipc_export_size_contract(pas):
    """include/pas_b200.h: PAS_IPC_EXPORT_BYTES is what pas_model_ipc_export writes per rank and what
    world.py all-gathers; the Python mirror and the header must agree."""
    text = open(HEADER).read()
    m = re.search(r"#define\s+PAS_IPC_EXPORT_BYTES\s+(\d+)", text)
    assert m and int(m.group(1)) == pas.model.IPC_EXPORT_BYTES

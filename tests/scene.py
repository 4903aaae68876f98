"""TEST INFRASTRUCTURE: the view / scene description consumed by tests/scene_render.py (the CUDA scene
kernel of tests/cuda/scene_kernel.cu) and by the oracle's renderer, with the reference's
integration-test scene as the canonical instance.

The scene is the one of atmosphere/reference/model_test.glsl: a sphere S resting on a spherical
planet P, lit by the sun and the sky, seen through the atmosphere with light shafts. Camera, view
matrix, exposure and tone map follow atmosphere/reference/model_test.cc:436-477, 726-731; the PSNR is
the reference's own formula (model_test.cc:750-765, which takes the square root of the mean square
error before the logarithm -- reproduced as written so that its thresholds can be quoted).
"""
from __future__ import annotations

import dataclasses
import math
from typing import Sequence

import numpy as np

# Albedo spectra of the test scene (model_test.cc:325-355): grass 360..800 nm (45 samples) and
# snow 360..420 nm (7 samples), constant outside their range, linear in between.
GRASS_ALBEDO = [
    0.018, 0.019, 0.019, 0.020, 0.022, 0.024, 0.027, 0.029, 0.030, 0.031, 0.032, 0.032, 0.032, 0.033,
    0.035, 0.040, 0.055, 0.073, 0.084, 0.089, 0.089, 0.079, 0.069, 0.063, 0.061, 0.057, 0.052, 0.051,
    0.048, 0.042, 0.039, 0.035, 0.035, 0.043, 0.087, 0.156, 0.234, 0.334, 0.437, 0.513, 0.553, 0.571,
    0.579, 0.581, 0.587]
GRASS_RANGE = (360.0, 800.0)
SNOW_ALBEDO = [0.796, 0.802, 0.807, 0.810, 0.818, 0.825, 0.826]
SNOW_RANGE = (360.0, 420.0)


def sample_uniform_spectrum(lo: float, hi: float, values: Sequence[float], wavelength: float) -> float:
    """Value at `wavelength` of a spectrum sampled uniformly on [lo, hi] (inclusive), constant
    outside and linear in between (dimensional_types scalar_function.h:84-98, 225-259)."""
    n = len(values)
    if wavelength <= lo:
        return values[0]
    if wavelength >= hi:
        return values[-1]
    u = (wavelength - lo) / (hi - lo) * (n - 1)
    i = min(int(math.floor(u)), n - 2)
    f = u - i
    return values[i] * (1.0 - f) + values[i + 1] * f


def grass_albedo(wavelength: float) -> float:
    return sample_uniform_spectrum(*GRASS_RANGE, GRASS_ALBEDO, wavelength)


def snow_albedo(wavelength: float) -> float:
    return sample_uniform_spectrum(*SNOW_RANGE, SNOW_ALBEDO, wavelength)


@dataclasses.dataclass
class SceneView:
    """Uniforms of the test-scene shader (model_test.cc:127-134), in the model's length unit."""
    camera: Sequence[float]
    earth_center: Sequence[float]
    sun_direction: Sequence[float]
    sun_size: Sequence[float]            # (tan, cos) of the sun's angular radius
    sphere_center: Sequence[float]
    sphere_radius: float
    model_from_clip: Sequence[float]     # 3x3 row major, clip (x, y, 1) -> world view ray
    ground_albedo: Sequence[float]       # at 680 / 550 / 440 nm
    sphere_albedo: Sequence[float]
    exposure: float
    use_luminance: bool
    width: int
    height: int


def model_from_clip(width: int, height: int, camera=(2.0, -8.0, 0.5), pitch: float = math.pi / 30.0,
                    fov_y: float = 50.0 / 180.0 * math.pi) -> np.ndarray:
    """model_test.cc:438-469; the reference builds it in float, so do we."""
    f = np.float32
    s, c = f(math.sin(f(pitch))), f(math.cos(f(pitch)))
    model_from_view = np.array([[1, 0, 0, camera[0]], [0, -s, -c, camera[1]], [0, c, -s, camera[2]],
                                [0, 0, 0, 1]], dtype=np.float32)
    tan_fov = f(math.tan(float(f(fov_y)) / 2.0))
    view_from_clip = np.array([[tan_fov * f(width) / f(height), 0, 0, 0], [0, tan_fov, 0, 0],
                               [0, 0, 0, -1], [0, 0, 1, 1]], dtype=np.float32)
    out = np.zeros((3, 3), dtype=np.float32)
    for row in range(3):
        for col in range(3):
            col2 = col if col < 2 else 3
            out[row, col] = (model_from_view[row, 0] * view_from_clip[0, col2] +
                             model_from_view[row, 1] * view_from_clip[1, col2] +
                             model_from_view[row, 2] * view_from_clip[2, col2])
    return out.astype(np.float64)


def model_test_view(sun_zenith_deg: float, sun_azimuth_deg: float, use_luminance: bool, *,
                    width: int = 640, height: int = 360, sun_angular_radius: float,
                    bottom_radius: float = 6360.0, length_unit_in_meters: float = 1000.0,
                    ground_albedo: Sequence[float] | None = None,
                    sphere_albedo: Sequence[float] | None = None) -> SceneView:
    """SetViewParameters + SetUp of the reference's integration test (model_test.cc:310-318, 436-477):
    camera at (2, -8, 0.5) km pitched by pi/30, 50 deg vertical field of view, sphere of radius 1 km
    resting on the ground at the origin, grass ground and snow sphere."""
    km = 1000.0 / length_unit_in_meters
    theta, phi = math.radians(sun_zenith_deg), math.radians(sun_azimuth_deg)
    lam = (680.0, 550.0, 440.0)
    return SceneView(
        camera=[2.0 * km, -8.0 * km, 0.5 * km],
        earth_center=[0.0, 0.0, -bottom_radius],
        sun_direction=[math.cos(phi) * math.sin(theta), math.sin(phi) * math.sin(theta), math.cos(theta)],
        sun_size=[math.tan(sun_angular_radius), math.cos(sun_angular_radius)],
        sphere_center=[0.0, 0.0, 1.0 * km], sphere_radius=1.0 * km,
        model_from_clip=list(model_from_clip(width, height, camera=(2.0 * km, -8.0 * km, 0.5 * km)).ravel()),
        ground_albedo=list(ground_albedo) if ground_albedo is not None else [grass_albedo(l) for l in lam],
        sphere_albedo=list(sphere_albedo) if sphere_albedo is not None else [snow_albedo(l) for l in lam],
        exposure=1e-4 if use_luminance else 10.0, use_luminance=use_luminance, width=width, height=height)


def tone_map(rgb: np.ndarray, exposure: float) -> np.ndarray:
    """[H, W, 3] radiance or luminance -> [H, W] ARGB words (model_test.cc:726-736)."""
    v = np.power(1.0 - np.exp(-np.asarray(rgb, dtype=np.float64) * exposure), 1.0 / 2.2)
    q = (v * 255.0).astype(np.uint32)
    return (np.uint32(255) << 24) | (q[..., 0] << 16) | (q[..., 1] << 8) | q[..., 2]


def psnr(image1: np.ndarray, image2: np.ndarray) -> float:
    """ComputePSNR of the reference, as written (model_test.cc:750-765)."""
    def channels(a):
        a = np.asarray(a, dtype=np.uint32)
        return np.stack([(a >> 16) & 0xFF, (a >> 8) & 0xFF, a & 0xFF], axis=-1).astype(np.float64)
    d = channels(image1) - channels(image2)
    square_error_sum = float((d * d).sum())
    if square_error_sum == 0.0:
        return float("inf")
    mean_square_error = math.sqrt(square_error_sum / (image1.shape[0] * image1.shape[1]))
    return 10.0 * math.log(255 * 255 / mean_square_error) / math.log(10.0)

"""Times the model_test.glsl scene kernel at 1920x1080 (device time, CUDA events inside the library)
for the table formats the reference supports. Usage under gpurun: python tools/render_bench.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import precomputed_atmospheric_scattering_b200 as pas  # noqa: E402
from tests import scene, scene_render  # noqa: E402

for n, combined, half in ((3, True, True), (3, False, True), (3, False, False), (15, True, True)):
    spec = pas.model_test_earth(n, combine_scattering_textures=combined, half_precision=half)
    model = pas.Model.from_spec(spec)
    model.Init(4)
    lum = n > 3
    for zen in (65.0, 88.0):
        view = scene.model_test_view(zen, 90.0, lum, width=1920, height=1080,
                                         sun_angular_radius=spec.sun_angular_radius)
        ms = []
        for _ in range(6):
            scene_render.render_scene(model, view)
            ms.append(scene_render.last_kernel_ms)
        print(f"wavelengths={n} combined={combined} half={half} zenith={zen}: 1080p scene kernel "
              f"{np.median(ms[1:]):.3f} ms ({1920 * 1080 / np.median(ms[1:]) / 1e6:.2f} Gpixel/s)")
    model.close()

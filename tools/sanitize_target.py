"""Short target for compute-sanitizer: one precompute of a small configuration that runs the SAME
kernel instantiations as the bench (row width 256 = 8 nu x 32 mu_s and a 256-wide transmittance table
-> the register-slot ray marches; 15 channels -> the packed-fp32 density kernel with 16-float texels,
orders 2 and 3), with registered host outputs (pipelined read-back, three streams) and one render.
Usage under gpurun:  compute-sanitizer --tool racecheck python tools/sanitize_target.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import precomputed_atmospheric_scattering_b200 as pas  # noqa: E402
from tests import scene, scene_render  # noqa: E402

SIZES = dict(transmittance_width=256, transmittance_height=8, scattering_r=4, scattering_mu=8,
             scattering_mu_s=32, scattering_nu=8, irradiance_width=64, irradiance_height=4)
spec = pas.small_planet()
spec.num_precomputed_wavelengths = 3 if "--rgb" in sys.argv else 15
spec.half_precision = True
model = pas.Model.from_spec(spec, sizes=SIZES)
host = {}
for w in (pas.TEXTURE_TRANSMITTANCE, pas.TEXTURE_SCATTERING, pas.TEXTURE_IRRADIANCE):
    i = model.texture_info(w)
    shape = ((i.depth,) if i.depth > 1 else ()) + (i.height, i.width, 4)
    host[w] = torch.zeros(shape, dtype=torch.float16 if i.bytes_per_channel == 2 else torch.float32).pin_memory().numpy()
model.set_host_outputs(transmittance=host[pas.TEXTURE_TRANSMITTANCE], scattering=host[pas.TEXTURE_SCATTERING],
                       irradiance=host[pas.TEXTURE_IRRADIANCE])
for _ in range(2):   # the second Init runs on recycled, dirty buffers
    model.Init(3)
S = model.texture(pas.TEXTURE_SCATTERING, as_float32=False)
assert np.array_equal(S, host[pas.TEXTURE_SCATTERING]) and np.isfinite(S.astype(np.float32)).all()
view = scene.model_test_view(65.0, 90.0, spec.num_precomputed_wavelengths > 3, width=64, height=36,
                                 sun_angular_radius=spec.sun_angular_radius)
rgb, _ = scene_render.render_scene(model, view)
assert np.isfinite(rgb).all()
print("sanitize target ok:", model.last_launch_count(), "launches,", {k: round(v, 3) for k, v in model.last_timings().items()})
model.close()
# a second model takes its buffers from the pool (initcheck: nothing may be read before it is written)
model = pas.Model.from_spec(spec, sizes=SIZES)
model.Init(2)
assert np.isfinite(model.scattering).all()
model.close()

"""Where the end-to-end time of bench.py goes: create / Init / read tables / destroy, host clock."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import precomputed_atmospheric_scattering_b200 as pas
spec = pas.earth(15, half_precision=True, combine_scattering_textures=True)
m = pas.Model.from_spec(spec); m.Init(4)
info = {w: m.texture_info(w) for w in (pas.TEXTURE_TRANSMITTANCE, pas.TEXTURE_SCATTERING, pas.TEXTURE_IRRADIANCE)}
host = {}
for w, i in info.items():
    shape = ((i.depth,) if i.depth > 1 else ()) + (i.height, i.width, 4)
    host[w] = torch.empty(shape, dtype=torch.float16 if i.bytes_per_channel == 2 else torch.float32).pin_memory().numpy()
m.close()
acc = {"create": 0.0, "init": 0.0, "read": 0.0, "close": 0.0}
N = 30
for it in range(N + 3):
    t0 = time.perf_counter(); m = pas.Model.from_spec(spec)
    t1 = time.perf_counter(); m.Init(4)
    t2 = time.perf_counter()
    for w, arr in host.items():
        m.texture(w, as_float32=False, out=arr)
    t3 = time.perf_counter(); m.close()
    t4 = time.perf_counter()
    if it >= 3:
        for k, v in zip(acc, (t1 - t0, t2 - t1, t3 - t2, t4 - t3)):
            acc[k] += v * 1e3 / N
print({k: round(v, 4) for k, v in acc.items()}, "total", round(sum(acc.values()), 4))

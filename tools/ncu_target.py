"""Short target for ncu: one warm-up and one profiled precompute of the bench workload (15
wavelengths, 4 orders) -- or RGB with --rgb. Usage under gpurun:
  ncu --set full --clock-control none --import-source on -k regex:'density|scattering' -s 12 -c 9 \
      -o gpurun_out/prof python tools/ncu_target.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import precomputed_atmospheric_scattering_b200 as pas  # noqa: E402

n = 3 if "--rgb" in sys.argv else 15
model = pas.Model.from_spec(pas.earth(n, half_precision=True))
for _ in range(2):
    model.Init(4)
print({k: round(v, 4) for k, v in model.last_timings().items()}, model.last_launch_count())
model.close()

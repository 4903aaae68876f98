"""Probe: e2e flow of bench.py at N > 1 with shared host tables; prints which layers mismatch."""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import precomputed_atmospheric_scattering_b200 as pas
from precomputed_atmospheric_scattering_b200 import world
rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
spec = pas.earth(15, half_precision=True)
ORDERS = int(os.environ.get("ORDERS", "4"))
m0 = pas.Model.from_spec(spec, device=local); world.attach(m0)
which = [pas.TEXTURE_TRANSMITTANCE, pas.TEXTURE_SCATTERING, pas.TEXTURE_IRRADIANCE]
shared = world.shared_host_tables(m0, which)
m0.Init(ORDERS)
want = {w: m0.texture(w, as_float32=False) for w in which}
m0.close()
for step in range(6):
    m = pas.Model.from_spec(spec, device=local); world.attach(m)
    m.set_host_outputs(transmittance=shared.arrays[which[0]], scattering=shared.arrays[which[1]], irradiance=shared.arrays[which[2]])
    m.set_host_output_mode(True)
    m.Init(ORDERS)
    got = {w: np.array(shared.arrays[w]) for w in which}
    own = {w: m.texture(w, as_float32=False) for w in which}
    m.close()
    for w in which:
        a, b = got[w].astype(np.float32), want[w].astype(np.float32)
        if not np.array_equal(a, b):
            if a.ndim == 4:
                bad = [k for k in range(a.shape[0]) if not np.array_equal(a[k], b[k])]
                print(f"rank {rank} step {step} table {w}: layers {bad} differ; own-vs-want equal: {np.array_equal(own[w], want[w])}; max rel {np.abs(a-b).max()/np.abs(b).max():.3e}", flush=True)
            else:
                print(f"rank {rank} step {step} table {w} differs: max {np.abs(a-b).max():.3e}; own-vs-want equal: {np.array_equal(own[w], want[w])}", flush=True)
print(f"rank {rank} done", flush=True)
dist.barrier(); shared.close(); dist.destroy_process_group()

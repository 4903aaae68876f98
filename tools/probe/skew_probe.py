"""Probe: how much of the first phase of a multi-GPU Init is start skew between the ranks' host threads?
Compares back-to-back Inits with Inits that start right after a host barrier."""
import os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import precomputed_atmospheric_scattering_b200 as pas
from precomputed_atmospheric_scattering_b200 import world
rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
m = pas.Model.from_spec(pas.earth(15, half_precision=True), device=local); world.attach(m)
for _ in range(5): m.Init(4)
for mode in ("back-to-back", "host barrier first"):
    acc, tot = {}, []
    for _ in range(20):
        if mode != "back-to-back":
            dist.barrier(); torch.cuda.synchronize()
        m.Init(4)
        tm = m.last_timings(); tot.append(sum(tm.values()))
        for k, v in tm.items(): acc[k] = acc.get(k, 0) + v / 20
    t = torch.tensor([np.mean(tot)], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(mode, "max over ranks", round(t.item(), 4), {k: round(v, 4) for k, v in acc.items()}, flush=True)
m.close(); dist.barrier(); dist.destroy_process_group()

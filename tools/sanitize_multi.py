"""compute-sanitizer target for the multi-GPU exchange: 2 ranks (one process per GPU, spawned here), the
small configuration of tools/sanitize_target.py (bench kernel instantiations), attached to a world with
the exchange named by PAS_EXCHANGE (symm: multicast stores + the 2-CTA cluster / distributed-shared-memory
output path of the density kernel; peer: unicast IPC mappings), two Inits, shared host tables.
Usage under gpurun --gpus 2:
  compute-sanitizer --tool memcheck --target-processes all python tools/sanitize_multi.py"""
import os
import socket
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
SIZES = dict(transmittance_width=256, transmittance_height=8, scattering_r=4, scattering_mu=8,
             scattering_mu_s=32, scattering_nu=8, irradiance_width=64, irradiance_height=4)


def worker(rank, world_size, port):
    import torch.distributed as dist
    import precomputed_atmospheric_scattering_b200 as pas
    from precomputed_atmospheric_scattering_b200 import world
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world_size, device_id=torch.device("cuda", rank))
    spec = pas.small_planet()
    spec.num_precomputed_wavelengths = 15
    spec.half_precision = True
    model = pas.Model.from_spec(spec, sizes=SIZES, device=rank)
    world.attach(model)
    which = [pas.TEXTURE_TRANSMITTANCE, pas.TEXTURE_SCATTERING, pas.TEXTURE_IRRADIANCE]
    shared = world.shared_host_tables(model, which, tag="sanitize")
    model.set_host_outputs(transmittance=shared.arrays[which[0]], scattering=shared.arrays[which[1]],
                           irradiance=shared.arrays[which[2]])
    model.set_host_output_mode(True)
    for _ in range(2):
        model.Init(3)
    S = model.texture(pas.TEXTURE_SCATTERING, as_float32=False)
    assert np.array_equal(S, shared.arrays[which[1]]) and np.isfinite(S.astype(np.float32)).all()
    print(f"rank {rank}: sanitize multi ok, exchange {os.environ.get('PAS_EXCHANGE', 'symm')}, "
          f"{model.last_launch_count()} launches", flush=True)
    model.set_host_outputs()
    shared.close()
    model.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(worker, args=(2, port), nprocs=2, join=True)

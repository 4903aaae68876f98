"""Multi-GPU parity: r-slab sharding over 2 (or more) B200s with the exchange inside the library --
stores fused into the density kernel (through the NVLS multicast address of a symmetric arena, or
unicast into CUDA IPC mappings) + flag barriers, or NCCL collectives -- must give
the single-GPU tables bit for bit (same kernels, same per-texel order of operations; the exchange
only moves data), except the irradiance, whose per-slab partial sums are added in a different order
(compared at 1e-6). Skipped with fewer than 2 GPUs."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

SIZES = dict(transmittance_width=64, transmittance_height=16, scattering_r=8, scattering_mu=32,
             scattering_mu_s=8, scattering_nu=8, irradiance_width=16, irradiance_height=4)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _spawn(fn, world_size, *args):
    """mp.spawn on a free rendezvous port; a port that another process grabbed in between (EADDRINUSE)
    is not a test failure: try another one."""
    import torch.multiprocessing as mp
    for attempt in range(4):
        try:
            mp.spawn(fn, args=(world_size, _free_port()) + args, nprocs=world_size, join=True)
            return
        except Exception as e:  # ProcessRaisedException carries the child's traceback as text
            if "EADDRINUSE" not in str(e) or attempt == 3:
                raise


def _worker(rank, world_size, port, out_dir, full_size, exchange):
    import torch.distributed as dist

    import precomputed_atmospheric_scattering_b200 as pas
    from precomputed_atmospheric_scattering_b200 import world
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world_size,
                            device_id=torch.device("cuda", rank))
    try:
        spec = pas.earth(15, half_precision=False) if full_size else pas.small_planet()
        kw = {} if full_size else dict(sizes=SIZES)
        for attempt in range(2):  # the second model re-uses the cached communicator
            model = pas.Model.from_spec(spec, device=rank, **kw)
            assert world.attach(model, exchange=exchange) == (rank, world_size)
            model.Init(4)
            np.savez(os.path.join(out_dir, f"rank{rank}_{attempt}.npz"), S=model.scattering,
                     E=model.irradiance, T=model.transmittance)
            model.close()
        assert pas.world_is_cached(rank, rank, world_size) == (exchange == "nccl")
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("exchange", ["symm", "peer", "nccl"])
@pytest.mark.parametrize("full_size", [False, True], ids=["small", "earth15"])
@pytest.mark.timeout(600)
def test_two_gpus_match_one(tmp_path, pas, full_size, exchange):
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world_size = int(os.environ.get("PAS_TEST_WORLD", "2"))
    if n < world_size:
        pytest.skip(f"needs >= {world_size} GPUs")
    _spawn(_worker, world_size, str(tmp_path), full_size, exchange)
    spec = pas.earth(15, half_precision=False) if full_size else pas.small_planet()
    single = pas.Model.from_spec(spec, device=0, **({} if full_size else dict(sizes=SIZES)))
    single.Init(4)
    S, E, T = single.scattering, single.irradiance, single.transmittance
    for rank in range(world_size):
        for attempt in range(2):
            got = np.load(os.path.join(tmp_path, f"rank{rank}_{attempt}.npz"))
            assert np.array_equal(got["T"], T)
            # orders >= 3 consume the all-reduced irradiance, whose summation order differs
            assert np.allclose(got["S"], S, rtol=2e-6, atol=1e-7 * np.abs(S).max())
            assert np.allclose(got["E"], E, rtol=1e-5, atol=1e-7 * np.abs(E).max())
    single.close()


def _variant_worker(rank, world_size, port, out_dir, exchange):
    import torch.distributed as dist

    import precomputed_atmospheric_scattering_b200 as pas
    from precomputed_atmospheric_scattering_b200 import world
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world_size,
                            device_id=torch.device("cuda", rank))
    try:
        for name, spec, orders, sizes in _variants(pas):
            model = pas.Model.from_spec(spec, device=rank, **sizes)
            world.attach(model, exchange=exchange)
            model.Init(orders)
            out = dict(S=model.scattering, E=model.irradiance, T=model.transmittance)
            if not spec.combine_scattering_textures:
                out["M"] = model.single_mie_scattering
            np.savez(os.path.join(out_dir, f"{name}_rank{rank}.npz"), **out)
            model.close()
        # shared host tables: every rank copies only the layers it computed into ONE set of host tables
        # (POSIX shared memory, page-locked by every rank); when Init returns on a rank, all of it is there
        for spec, kw in ((pas.small_planet(), dict(sizes=SIZES)), (pas.earth(15, half_precision=True), {})):
            m = pas.Model.from_spec(spec, device=rank, **kw)
            world.attach(m, exchange=exchange)
            which = [pas.TEXTURE_TRANSMITTANCE, pas.TEXTURE_SCATTERING, pas.TEXTURE_IRRADIANCE]
            shared = world.shared_host_tables(m, which, tag=f"test{len(kw)}")
            m.set_host_outputs(transmittance=shared.arrays[pas.TEXTURE_TRANSMITTANCE],
                               scattering=shared.arrays[pas.TEXTURE_SCATTERING],
                               irradiance=shared.arrays[pas.TEXTURE_IRRADIANCE])
            m.set_host_output_mode(True)
            plain = pas.Model.from_spec(spec, device=rank, **kw)     # every rank copies everything, after Init
            world.attach(plain, exchange=exchange)
            plain.Init(4)
            want = {w: plain.texture(w, as_float32=False) for w in which}
            plain.close()
            for attempt in range(2):
                dist.barrier()
                if rank == 0:
                    for a in shared.arrays.values():
                        a[...] = 0
                dist.barrier()
                m.Init(4)
                for w in which:
                    assert np.array_equal(shared.arrays[w], m.texture(w, as_float32=False)), (w, attempt, rank)
                    assert np.array_equal(shared.arrays[w], want[w]), (w, attempt, rank)
            m.set_host_outputs()
            shared.close()
            m.close()
        # one barrier sequence per peer world: a second model of the world cannot start its Init while
        # the first is in flight (its barriers would release the first model's early); once the first
        # has been waited for, it runs and gives the same tables
        a = pas.Model.from_spec(pas.small_planet(), device=rank, sizes=SIZES)
        b = pas.Model.from_spec(pas.small_planet(), device=rank, sizes=SIZES)
        world.attach(a, exchange=exchange)
        world.attach(b, exchange=exchange)
        a.InitAsync(3)
        try:
            b.InitAsync(3)
            raise AssertionError("two models of one peer world were in flight")
        except pas.PasError as e:
            assert e.status == 5, e     # PAS_ERR_STATE
        a.Wait()
        Sa, Ea = a.scattering, a.irradiance
        b.Init(3)
        # (the product tables are each model's own, also when the models share a symmetric arena)
        assert np.array_equal(Sa, b.scattering) and np.array_equal(Ea, b.irradiance)
        assert np.array_equal(Sa, a.scattering)
        a.close()
        b.close()
        dist.barrier()
    finally:
        dist.destroy_process_group()


def _variants(pas):
    """Paths of the peer exchange the Earth runs above do not take: a separate single-Mie table with
    fp16 products (second final push), 24 channels = two launch groups (the buffer parity carries
    over from one group to the next; at the reference's full sizes too, where a rank that finishes a
    group early would otherwise overwrite the transmittance rows its peer still ray-marches with),
    and Init(1) (no order loop: only the final exchange), also with two groups (no barrier at all
    between the groups but the one in front of the transmittance rows)."""
    sep = pas.small_planet()
    sep.combine_scattering_textures, sep.half_precision = False, True
    wide = pas.small_planet()
    wide.num_precomputed_wavelengths = 24
    small = dict(sizes=SIZES)
    return [("separate_mie_fp16", sep, 3, small), ("two_groups", wide, 3, small),
            ("single_order", pas.small_planet(), 1, small),
            ("two_groups_full_size", pas.earth(24, half_precision=False), 2, {}),
            ("two_groups_full_size_single_order", pas.earth(24, half_precision=False), 1, {})]


@pytest.mark.parametrize("exchange", ["symm", "peer"])
@pytest.mark.timeout(600)
def test_peer_exchange_variants(tmp_path, pas, exchange):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    _spawn(_variant_worker, 2, str(tmp_path), exchange)
    for name, spec, orders, sizes in _variants(pas):
        single = pas.Model.from_spec(spec, device=0, **sizes)
        single.Init(orders)
        want = dict(S=single.scattering, E=single.irradiance, T=single.transmittance)
        if not spec.combine_scattering_textures:
            want["M"] = single.single_mie_scattering
        for rank in range(2):
            got = np.load(os.path.join(tmp_path, f"{name}_rank{rank}.npz"))
            for key, ref in want.items():
                ref = np.asarray(ref, dtype=np.float64)
                tol = 2e-3 if spec.half_precision and key in ("S", "M") else 2e-6
                assert np.allclose(np.asarray(got[key], dtype=np.float64), ref, rtol=tol,
                                   atol=1e-6 * np.abs(ref).max()), (name, rank, key)
        single.close()

"""world_size-2 test of the multi-GPU host logic on CPU (gloo): unique-id rendezvous, r-slab
ownership, max-over-ranks timing, and the sharding contract itself -- the scattering-density pass
of layer k reads only layer k of the previous order, so r-slabs computed by
different ranks and all-gathered reproduce the single-process table. The per-slab
compute here is the CPU oracle standing in for the kernels (checker only); the GPU version of this
test is tests/test_gpu_multi.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import precomputed_atmospheric_scattering_b200 as pas
from precomputed_atmospheric_scattering_b200 import world

SIZES = dict(t_w=32, t_h=8, r=4, mu=8, mu_s=4, nu=4, e_w=8, e_h=4)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world_size, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        from oracle import oracle as orc
        # 1. every rank ends up with rank 0's unique id
        uid = world.broadcast_unique_id(lambda: bytes(range(128)))
        assert uid == bytes(range(128))
        # 1b. the IPC-handle exchange of the peer path: every rank gets every blob, in rank order
        blobs = world.all_gather_bytes(bytes([rank]) * 400)
        assert blobs == b"".join(bytes([r]) * 400 for r in range(world_size))
        # 1c. world.attach over the peer exchange: every rank exports, the blobs are all-gathered in
        # rank order and handed to attach_peers; the NCCL exchange broadcasts rank 0's unique id
        class FakeModel:
            device = 0

            def __init__(self):
                self.calls = []

            def ipc_export(self, r, w):
                self.calls.append(("export", r, w))
                return bytes([10 + r]) * 464

            def attach_peers(self, blob, per_rank):
                self.calls.append(("peers", blob, per_rank))

            def attach_world(self, r, w, uid):
                self.calls.append(("world", r, w, uid))

            def exchange_bytes(self, w):
                return 3 << 20

            def attach_symmetric(self, r, w, ptrs, mc, nbytes):
                self.calls.append(("symm", r, w, tuple(ptrs), mc, nbytes))

        # 1b'. the symmetric exchange: the arena is allocated once per (rank, world) -- a collective --
        # and every later model attaches to it without one; a rank that cannot allocate takes every
        # rank to the next exchange together
        allocations = []

        def fake_arena(nbytes, group=None):
            allocations.append(nbytes)
            return world.Arena([0x1000 * (r + 1) for r in range(world_size)], 0xabc000, nbytes)

        real_alloc, world.allocate_arena = world.allocate_arena, fake_arena
        try:
            fm = FakeModel()
            assert world.attach(fm, exchange="symm") == (rank, world_size)
            assert fm.calls == [("symm", rank, world_size, tuple(0x1000 * (r + 1) for r in range(world_size)),
                                 0xabc000, allocations[0])] and allocations[0] >= 3 << 20
            fm = FakeModel()
            assert world.attach(fm) == (rank, world_size) and len(allocations) == 1   # cached arena, default exchange
            assert fm.calls[0][0] == "symm"

            class Bigger(FakeModel):
                def exchange_bytes(self, w):
                    return 64 << 20

            assert world.attach(Bigger()) == (rank, world_size) and len(allocations) == 2 and allocations[1] >= 64 << 20
        finally:
            world.allocate_arena = real_alloc
            world._ARENAS.clear()
        # no CUDA here: the real allocator fails on every rank, the default exchange falls through to peer
        import warnings as _w
        fm = FakeModel()
        with _w.catch_warnings():
            _w.simplefilter("ignore")
            assert world.attach(fm) == (rank, world_size)
        assert fm.calls[0][0] == "export" and (rank, world_size) in world._NO_SYMM
        try:
            world.attach(FakeModel(), exchange="symm")
            raise AssertionError("explicit symmetric exchange fell back")
        except RuntimeError:
            pass
        world._PEER_WORLDS.clear()

        fm = FakeModel()
        assert world.attach(fm, exchange="peer") == (rank, world_size)
        assert fm.calls[0] == ("export", rank, world_size)
        assert fm.calls[1] == ("peers", b"".join(bytes([10 + r]) * 464 for r in range(world_size)), 464)
        assert (rank, world_size) in world._PEER_WORLDS      # later attaches skip the agreement round
        # ... and exchange the blobs with one collective, whose status byte takes every rank out
        # together when one rank's export fails (instead of leaving the others in the all-gather)
        fm = FakeModel()
        assert world.attach(fm, exchange="peer") == (rank, world_size)
        assert fm.calls[1] == ("peers", b"".join(bytes([10 + r]) * 464 for r in range(world_size)), 464)

        class ExportFails(FakeModel):
            def ipc_export(self, r, w):
                if r == 1:
                    raise MemoryError("out of memory growing the peer buffers")
                return super().ipc_export(r, w)

        fm = ExportFails()
        try:
            world.attach(fm, exchange="peer")
            raise AssertionError("a failed export went unnoticed")
        except (MemoryError, RuntimeError) as e:
            assert isinstance(e, MemoryError) == (rank == 1)
        assert not any(c[0] == "peers" for c in fm.calls)
        world._PEER_WORLDS.clear()
        # a rank that cannot map peer memory drags every rank to the NCCL exchange, together
        class NoPeers(FakeModel):
            def attach_peers(self, blob, per_rank):
                if rank == 1:
                    raise RuntimeError("no peer access")
                super().attach_peers(blob, per_rank)

        import precomputed_atmospheric_scattering_b200.model as model_mod
        saved = (model_mod.world_is_cached, model_mod.nccl_unique_id)
        model_mod.world_is_cached = lambda d, r, w: False
        model_mod.nccl_unique_id = lambda: bytes(range(128))
        try:
            import warnings
            fm = NoPeers()
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                assert world.attach(fm) == (rank, world_size)
            assert fm.calls[-1] == ("world", rank, world_size, bytes(range(128)))
            try:
                world.attach(NoPeers(), exchange="peer")      # asked for explicitly: no silent fallback
                raise AssertionError("explicit peer exchange fell back")
            except RuntimeError:
                pass
        finally:
            model_mod.world_is_cached, model_mod.nccl_unique_id = saved
        try:
            world.attach(FakeModel(), exchange="carrier pigeon")
            raise AssertionError("unknown exchange accepted")
        except ValueError:
            pass
        # 2. the slowest rank defines the step time
        assert world.max_over_ranks(10.0 + rank) == 10.0 + world_size - 1
        # 3. r-slab sharded density pass == single-process pass, bit for bit
        spec = pas.small_planet()
        cp = pas.channel_params(spec, [680.0, 550.0, 440.0])
        sz = orc.Sizes(**SIZES)
        o = orc.Oracle(cp, sz)
        T = o.transmittance()
        dE = o.direct_irradiance(T)
        dR, dM = o.single_scattering(T)
        k0, k1 = world.slab(sz.r, rank, world_size)
        # this rank only holds its own layers of the previous order
        mine = lambda X: np.where((np.arange(sz.r) >= k0)[None, :, None, None] &
                                  (np.arange(sz.r) < k1)[None, :, None, None], X, 0.0)
        dJ = o.scattering_density(T, mine(dR), mine(dM), np.zeros_like(dR), dE, 2,
                                  rows=(k0 * sz.mu, k1 * sz.mu))
        part = torch.from_numpy(np.ascontiguousarray(dJ[:, k0:k1]))
        parts = [torch.empty_like(part) for _ in range(world_size)]
        dist.all_gather(parts, part)
        gathered = np.concatenate([p.numpy() for p in parts], axis=1)
        full = o.scattering_density(T, dR, dM, np.zeros_like(dR), dE, 2)
        # the literal 4-D lookup of the oracle touches the neighbouring layer with a weight of
        # ~1e-12; the kernels use layer k exactly
        assert np.allclose(gathered, full, rtol=1e-9, atol=1e-300)
        # 4. irradiance partial sums: all-reduce(sum) of per-slab contributions == full integral
        partial = o.indirect_irradiance(mine(dR), mine(dM), np.zeros_like(dR), 1)
        t = torch.from_numpy(partial.copy())
        dist.all_reduce(t)
        want = o.indirect_irradiance(dR, dM, np.zeros_like(dR), 1)
        assert np.allclose(t.numpy(), want, rtol=1e-13, atol=1e-300)
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_world(tmp_path, orc):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert sorted(os.listdir(tmp_path)) == ["ok0", "ok1"]

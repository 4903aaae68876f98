"""Multi-GPU plumbing: one process per GPU, rendezvous through ``torch.distributed``.

The compute and the NVLink exchange live in libpas_b200.so (r-slab sharding of every 3-D pass). Three
exchanges are built (include/pas_b200.h), tried in this order:
  * ``symm`` (default): a symmetric arena per rank (torch.distributed._symmetric_memory: CUDA VMM
    allocations mapped by every rank, plus -- on NVSwitch boxes -- one NVLS multicast address for all
    of them), allocated ONCE per (rank, world) and shared by every model attached afterwards, so that
    attaching a model costs no collective. The scattering-density kernel stores its r-slab through
    the multicast address: one store reaches every rank. Flag barriers in device memory;
  * ``peer``: the same kernels over CUDA IPC mappings of each model's own buffers (one all-gather of
    464 bytes per rank and model), unicast stores to every peer;
  * ``nccl``: all-gather of the density slabs / all-reduce of the irradiance partial sums.
What is left for the host is to set up those mappings (or an NCCL unique id) and to agree on who
owns which r-layers; that is all this module does. It works with any torch.distributed backend
(``nccl`` on the GPU box, ``gloo`` in the CPU tests).
"""
from __future__ import annotations

import os
import warnings
from typing import Callable, List, Optional, Tuple

import torch
import torch.distributed as dist

UNIQUE_ID_BYTES = 128  # PAS_NCCL_UNIQUE_ID_BYTES
_PEER_WORLDS = set()   # (rank, world) for which this process has agreed on the peer exchange
_ARENAS = {}           # (rank, world) -> Arena: the symmetric arena of this process in that world
_NO_SYMM = set()       # (rank, world) for which the symmetric exchange was tried and is unavailable
_RETIRED_ARENAS = []   # arenas replaced by larger ones; never unmapped while the process lives


class Arena:
    """A symmetric arena: ``ptrs[r]`` = rank r's arena mapped in this process, ``multicast`` = the
    NVLS multicast address of all of them (0 = none), ``keep`` = whatever owns the memory."""

    def __init__(self, ptrs, multicast, nbytes, keep=None):
        self.ptrs, self.multicast, self.nbytes, self.keep = list(ptrs), int(multicast or 0), int(nbytes), keep


def allocate_arena(nbytes: int, group=None) -> Arena:
    """COLLECTIVE: allocates ``nbytes`` of zeroed symmetric memory on every rank and maps every rank's
    allocation into every process (torch.distributed._symmetric_memory). Raises where the platform has
    no symmetric memory (no peer access, no fabric handles, a CPU-only process group)."""
    import torch.distributed._symmetric_memory as symm_mem
    device = torch.device("cuda", torch.cuda.current_device())
    group = group if group is not None else dist.group.WORLD
    t = symm_mem.empty(int(nbytes), dtype=torch.uint8, device=device)
    h = symm_mem.rendezvous(t, group)
    t.zero_()
    torch.cuda.synchronize()
    dist.barrier(group)      # nobody signals a flag word before every arena is zero
    multicast = h.multicast_ptr if os.environ.get("PAS_MULTICAST", "1") != "0" else 0
    return Arena(h.buffer_ptrs, multicast, nbytes, keep=(t, h))


def slab(r_n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous r-layers [k_begin, k_end) owned by ``rank`` (same rule as pas_model::slab in
    csrc/pas_model.cu): layers are dealt in order, the first ``r_n % world`` ranks get one more."""
    if not (0 <= rank < world):
        raise ValueError("bad rank / world size")
    base, extra = divmod(r_n, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def slabs(r_n: int, world: int) -> List[Tuple[int, int]]:
    return [slab(r_n, r, world) for r in range(world)]


def supported_world(r_n: int, world: int) -> bool:
    """The NCCL all-gather moves equal slabs, so the layer count must divide evenly."""
    return world >= 1 and r_n % world == 0


def broadcast_unique_id(make_id: Callable[[], bytes], group=None, device: Optional[torch.device] = None) -> bytes:
    """Rank 0 creates the NCCL unique id (``nccl_unique_id`` of the C ABI), every rank returns the
    same 128 bytes."""
    rank = dist.get_rank(group)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" \
            else torch.device("cpu")
    buf = torch.zeros(UNIQUE_ID_BYTES, dtype=torch.uint8, device=device)
    if rank == 0:
        raw = make_id()
        if len(raw) != UNIQUE_ID_BYTES:
            raise ValueError("unique id must be 128 bytes")
        buf.copy_(torch.frombuffer(bytearray(raw), dtype=torch.uint8))
    dist.broadcast(buf, src=0, group=group)
    return bytes(buf.cpu().tolist())


def _host_device(group) -> torch.device:
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" \
        else torch.device("cpu")


def all_gather_bytes(blob: bytes, group=None, device: Optional[torch.device] = None) -> bytes:
    """Concatenation, in rank order, of every rank's ``blob`` (equal lengths)."""
    world = dist.get_world_size(group)
    device = device or _host_device(group)
    mine = torch.frombuffer(bytearray(blob), dtype=torch.uint8).to(device)
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine, group=group)
    return b"".join(bytes(p.cpu().numpy().tobytes()) for p in parts)


def attach(model, group=None, exchange: Optional[str] = None) -> Tuple[int, int]:
    """Attaches ``model`` (model.Model) to the default process group: returns (rank, world).
    ``exchange``: "symm" (default; env PAS_EXCHANGE overrides), "peer" or "nccl"; an exchange that the
    box cannot provide falls through to the next one on every rank together, unless it was asked for
    explicitly."""
    from .model import nccl_unique_id, world_is_cached
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if world == 1:
        model.attach_world(0, 1, None)
        return rank, world
    requested = exchange or os.environ.get("PAS_EXCHANGE")
    exchange = requested or "symm"
    if exchange == "symm":
        if (rank, world) not in _NO_SYMM:
            need = model.exchange_bytes(world)
            arena = _ARENAS.get((rank, world))
            if arena is None or arena.nbytes < need:
                # every rank creates the same models in the same order, so every rank is here together:
                # allocate (with headroom, so that a slightly larger model does not allocate again) and
                # agree that it worked everywhere
                arena, error = None, None
                try:
                    arena = allocate_arena(max(need + need // 8, 1 << 20), group)
                except Exception as e:
                    error = e
                if _all_ok(error is None, group):
                    if (rank, world) in _ARENAS:
                        # models attached earlier keep pointing into the smaller arena: it stays mapped
                        _RETIRED_ARENAS.append(_ARENAS[(rank, world)])
                    _ARENAS[(rank, world)] = arena
                else:
                    arena = None
                    _NO_SYMM.add((rank, world))
                    if requested == "symm":
                        raise error if error is not None else RuntimeError(
                            "another rank could not allocate its symmetric arena")
                    warnings.warn(f"symmetric-memory exchange unavailable ({error or 'on another rank'}): "
                                  "using CUDA IPC peer mappings")
            if arena is not None:
                # no collective from here on: the arena is mapped, the model only takes its tables in it
                model.attach_symmetric(rank, world, arena.ptrs, arena.multicast, arena.nbytes)
                return rank, world
        elif requested == "symm":
            raise RuntimeError("the symmetric-memory exchange is unavailable in this world")
        exchange = "peer"
    if exchange == "peer":
        # Every rank must end up on the same exchange: a rank whose GPU cannot export or map peer
        # memory (no P2P between the devices, IPC disabled in the container) tells the others, and
        # unless the peer exchange was asked for explicitly all of them fall back to NCCL together.
        if (rank, world) in _PEER_WORLDS:
            # this process has already mapped its peers once (and so has every other rank): the
            # exchange is settled, a failure now is an error. The one collective still carries a
            # status byte per rank, so that a rank whose export failed (out of memory while growing
            # its buffers) takes every rank out together instead of leaving them in the all-gather.
            from .model import IPC_EXPORT_BYTES
            blob, error = bytes(IPC_EXPORT_BYTES), None
            try:
                blob = model.ipc_export(rank, world)
            except Exception as e:  # PasError
                error = e
            packed = all_gather_bytes(bytes([0 if error else 1]) + blob, group)
            n = 1 + len(blob)
            ok = [packed[r * n] == 1 for r in range(world)]
            if not all(ok):
                raise error if error is not None else RuntimeError(
                    f"rank {ok.index(False)} could not export its peer tables")
            # (a rank failing to MAP a peer here raises alone; the others fail their next Init at the
            # first flag barrier, which times out and poisons every rank's flags: no silent corruption)
            model.attach_peers(b"".join(packed[r * n + 1:(r + 1) * n] for r in range(world)), len(blob))
            return rank, world
        blob, error = b"", None
        try:
            blob = model.ipc_export(rank, world)
        except Exception as e:  # PasError
            error = e
        if _all_ok(error is None, group):
            try:
                model.attach_peers(all_gather_bytes(blob, group), len(blob))
            except Exception as e:
                error = e
            if _all_ok(error is None, group):
                _PEER_WORLDS.add((rank, world))
                return rank, world
        if requested == "peer":
            raise error if error is not None else RuntimeError("another rank could not set up the peer exchange")
        warnings.warn(f"peer-memory exchange unavailable ({error or 'on another rank'}): using NCCL")
    elif exchange != "nccl":
        raise ValueError("exchange must be 'symm', 'peer' or 'nccl'")
    device = model.device if model.device is not None else torch.cuda.current_device()
    # the library keeps one communicator per (device, rank, world) for the life of the process;
    # every rank takes the same branch because every rank has attached the same number of models
    uid = None if world_is_cached(device, rank, world) else broadcast_unique_id(nccl_unique_id, group)
    model.attach_world(rank, world, uid)
    return rank, world


def _all_ok(ok: bool, group=None) -> bool:
    """True when ``ok`` holds on every rank (all-reduce MIN of a flag)."""
    t = torch.tensor([1 if ok else 0], dtype=torch.int32, device=_host_device(group))
    dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
    return bool(t.item())


def max_over_ranks(value: float, group=None) -> float:
    """Timing rule of the bench: the slowest rank defines the step."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return float(value)
    device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" \
        else torch.device("cpu")
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def sum_over_ranks(value: float, group=None) -> float:
    """Sum of ``value`` over the ranks (used by the bench to count the ranks whose parity check failed)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=_host_device(group))
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return float(t.item())


class SharedHostTables:
    """One set of host product tables for all the ranks of a box: POSIX shared memory (/dev/shm) mapped
    by every rank and page-locked for CUDA by each of them (cudaHostRegister), so that every rank can
    copy the layers it computed straight into the common tables (Model.set_host_output_mode(True)).
    ``arrays[which]`` are numpy views shaped like the model's tables."""

    def __init__(self, model, which, group=None, tag: str = "tables"):
        import numpy as np
        rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.arrays, self._maps, self._registered, self._paths = {}, [], [], []
        # a name every rank derives alike and that other jobs on the box do not share
        job = os.environ.get("MASTER_PORT", "0") + "_" + os.environ.get("TORCHELASTIC_RUN_ID", str(os.getppid()))
        for w in which:
            info = model.texture_info(w)
            shape = ((info.depth,) if info.depth > 1 else ()) + (info.height, info.width, 4)
            dtype = np.float16 if info.bytes_per_channel == 2 else np.float32
            nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
            path = f"/dev/shm/pas_b200_{job}_{tag}_{w}"
            if rank == 0:
                with open(path, "wb") as f:
                    f.truncate(nbytes)
            self._paths.append(path)
            self.arrays[w] = (path, shape, dtype, nbytes)
        if dist.is_initialized():
            dist.barrier(group)
        cudart = torch.cuda.cudart()
        for w, (path, shape, dtype, nbytes) in list(self.arrays.items()):
            a = np.memmap(path, dtype=dtype, mode="r+", shape=shape)
            err = cudart.cudaHostRegister(a.ctypes.data, nbytes, 0)
            if int(err) != 0:
                raise RuntimeError(f"cudaHostRegister({path}) failed: {err}")
            self._maps.append(a)
            self._registered.append(a.ctypes.data)
            self.arrays[w] = a
        self._rank, self._group = rank, group

    def close(self):
        cudart = torch.cuda.cudart()
        for p in self._registered:
            cudart.cudaHostUnregister(p)
        self._registered = []
        self.arrays, self._maps = {}, []
        if dist.is_initialized():
            dist.barrier(self._group)
        if self._rank == 0:
            for path in self._paths:
                try:
                    os.unlink(path)
                except OSError:
                    pass


def shared_host_tables(model, which, group=None, tag: str = "tables") -> SharedHostTables:
    """COLLECTIVE: host tables shared by the ranks (see SharedHostTables)."""
    return SharedHostTables(model, which, group, tag)

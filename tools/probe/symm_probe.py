"""Probe: does torch's symmetric memory (CUDA VMM + multicast / NVLS) work on this box?
torchrun --nproc-per-node N tools/probe/symm_probe.py"""
import os
import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
try:
    t = symm_mem.empty(64 * 1024 * 1024 // 4, dtype=torch.float32, device=torch.device("cuda", local))
    h = symm_mem.rendezvous(t, dist.group.WORLD)
    print(f"rank {rank}: backend={symm_mem.get_backend(torch.device('cuda', local)) if hasattr(symm_mem, 'get_backend') else '?'} "
          f"multicast_support={h.has_multicast_support} multicast_ptr={h.multicast_ptr:#x} "
          f"buffer_ptrs={[hex(p) for p in h.buffer_ptrs]} signal_pads={len(h.signal_pad_ptrs)} size={h.buffer_size}", flush=True)
    # peer write test through buffer_ptrs: rank r writes r+1 into slot r of every rank
    t.zero_()
    h.barrier()
    for p in range(world):
        peer = h.get_buffer(p, (world,), torch.float32)
        peer[rank] = float(rank + 1)
    h.barrier()
    torch.cuda.synchronize()
    print(f"rank {rank}: gathered {t[:world].tolist()}", flush=True)
except Exception as e:
    import traceback; traceback.print_exc()
    print(f"rank {rank}: FAILED {type(e).__name__}: {e}", flush=True)
dist.barrier()
dist.destroy_process_group()

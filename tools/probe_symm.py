"""Does torch's symmetric memory give peer-mapped pointers on this box?  (exploration for a fused pass + exchange)"""
import os, sys, time
import torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
try:
    import torch.distributed._symmetric_memory as symm
    t = symm.empty(1 << 20, dtype=torch.float64, device=f"cuda:{local}")
    h = symm.rendezvous(t, dist.group.WORLD.group_name)
    t.fill_(rank + 1.0)
    dist.barrier(); torch.cuda.synchronize()
    peer = h.get_buffer((rank + 1) % world, (1 << 20,), torch.float64)
    v = float(peer[123].item())
    n = 1 << 26
    big = symm.empty(n, dtype=torch.float64, device=f"cuda:{local}")
    hb = symm.rendezvous(big, dist.group.WORLD.group_name)
    pb = hb.get_buffer((rank + 1) % world, (n,), torch.float64)
    dst = torch.empty(n, dtype=torch.float64, device=f"cuda:{local}")
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); dst.copy_(pb); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"rank {rank}: peer value {v} (expect {(rank + 1) % world + 1}), ptrs {[hex(p) for p in h.buffer_ptrs][:2]}, P2P read {n * 8 / ms / 1e6:.1f} GB/s", flush=True)
except Exception as e:
    print(f"rank {rank}: symmetric memory failed: {e!r}", flush=True)
dist.barrier()
dist.destroy_process_group()

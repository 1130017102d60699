"""Probe: NCCL all-gather bandwidth between the ranks of this box + transports NCCL picked."""
import os, time, torch, torch.distributed as dist
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
for mb in (16, 160):
    n = mb * 1024 * 1024 // 8
    x = torch.ones(n, dtype=torch.float64, device="cuda") * rank
    out = torch.empty(world * n, dtype=torch.float64, device="cuda")
    for _ in range(3):
        dist.all_gather_into_tensor(out, x)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        dist.all_gather_into_tensor(out, x)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    if rank == 0:
        print(f"all_gather {mb} MB/rank x {world}: {ms:.3f} ms  -> {mb*(world-1)/ms:.1f} GB/s per rank in", flush=True)
# peer access
if rank == 0:
    for j in range(torch.cuda.device_count()):
        print("can_device_access_peer(0,%d) =" % j, torch.cuda.can_device_access_peer(0, j) if j else "-")
# raw peer copy via torch (single process view)
if rank == 0 and torch.cuda.device_count() > 1:
    a = torch.ones(160 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda:0")
    b = torch.empty_like(a, device="cuda:1")
    for _ in range(2): b.copy_(a)
    torch.cuda.synchronize(0); torch.cuda.synchronize(1)
    t0 = time.perf_counter()
    for _ in range(5): b.copy_(a)
    torch.cuda.synchronize(0); torch.cuda.synchronize(1)
    print(f"peer copy 160 MB: {(time.perf_counter()-t0)/5*1e3:.3f} ms")
dist.barrier()
dist.destroy_process_group()

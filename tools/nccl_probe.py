"""Latency of small NCCL all_gathers on this box (run under torchrun)."""
import os, time, torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
for nbytes in (256, 16384, 180224, 4 << 20):
    x = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
    out = torch.zeros(nbytes * world, dtype=torch.uint8, device=dev)
    for _ in range(10):
        dist.all_gather_into_tensor(out, x)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(100):
        dist.all_gather_into_tensor(out, x)
    e1.record(); torch.cuda.synchronize(); t1 = time.perf_counter()
    if rank == 0:
        print(f"all_gather {nbytes:8d} B/rank: {e0.elapsed_time(e1)*10:.1f} us/call (events), {(t1-t0)*1e4:.1f} us/call (wall)")
can = torch.cuda.can_device_access_peer(local, (local + 1) % world)
if rank == 0:
    print("peer access:", can)
dist.destroy_process_group()

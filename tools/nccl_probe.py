"""NCCL all_gather_into_tensor latency by message size (torchrun, N ranks) — DESIGN.md 6: 38 us for 8.3 MB/rank at N = 2."""
import os, torch, torch.distributed as dist
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
dev = torch.device("cuda")
def bench(fn, n=30):
    for _ in range(5): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for mb in (0.01, 0.25, 1.0, 2.4, 8.3, 32.0):
    n = int(mb * 1e6) // 16 * 16
    mine = torch.empty(n, dtype=torch.uint8, device=dev)
    out = torch.empty(world * n, dtype=torch.uint8, device=dev)
    t = bench(lambda: dist.all_gather_into_tensor(out, mine))
    if rank == 0: print(f"all_gather {mb:6.2f} MB/rank x{world}: {t*1e3:8.1f} us", flush=True)
dist.destroy_process_group()

"""Pinned host<->device copy bandwidth on the GPU box (the e2e step uploads 39 MB; ~52 GB/s measured)."""
import torch, time
dev = torch.device("cuda:0")
h = torch.randn(64, 3, 224, 224).pin_memory()
d = torch.empty_like(h, device=dev)
hb = torch.empty(8 << 20, dtype=torch.uint8).pin_memory()
db = torch.empty(8 << 20, dtype=torch.uint8, device=dev)
for name, src, dst in (("H2D 38.5MB", h, d), ("D2H 8MB", db, hb), ("H2D 8MB", hb, db)):
    for _ in range(3):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        dst.copy_(src, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"{name}: {ms:.3f} ms  {src.numel() * src.element_size() / ms / 1e6:.1f} GB/s")

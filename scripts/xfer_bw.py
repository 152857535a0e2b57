"""PCIe ceilings of the box: pinned <-> device copy bandwidth (what bounds the e2e path)."""
import torch, time
dev = torch.device("cuda:0")
for mb in (16, 128, 512):
    n = mb << 20
    h = torch.empty(n, dtype=torch.uint8).pin_memory(); d = torch.empty(n, dtype=torch.uint8, device=dev)
    for name, fn in (("H2D", lambda: d.copy_(h, non_blocking=True)), ("D2H", lambda: h.copy_(d, non_blocking=True))):
        for _ in range(2): fn()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(5): fn()
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
        print(f"{name} {mb:4d} MiB pinned: {n / dt / 1e9:6.1f} GB/s")

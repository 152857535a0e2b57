"""Host<->device copy bandwidth of pinned buffers allocated on each NUMA node (tuning aid for the Arrow bridge's staging).
The allocating thread is pinned to one node's CPUs before cudaHostAlloc, so the pages land there (first touch)."""
import glob
import os
import subprocess

import torch

dev = torch.device("cuda:0")
bdf = torch.cuda.get_device_properties(0).pci_bus_id if hasattr(torch.cuda.get_device_properties(0), "pci_bus_id") else None
try:
    out = subprocess.run(["nvidia-smi", "--query-gpu=index,pci.bus_id", "--format=csv,noheader"], capture_output=True, text=True).stdout
    print("gpus:", out.strip().replace("\n", " | "))
    bus = out.strip().split("\n")[0].split(",")[1].strip().lower()
    bus = bus[4:] if bus.startswith("0000") and len(bus) > 12 else bus
    p = f"/sys/bus/pci/devices/{bus}"
    print("numa_node", open(p + "/numa_node").read().strip(), "local_cpulist", open(p + "/local_cpulist").read().strip())
except Exception as e:  # noqa
    print("topology probe failed", e)
nodes = sorted(glob.glob("/sys/devices/system/node/node[0-9]*"))
print("nodes:", [(os.path.basename(n), open(n + "/cpulist").read().strip()) for n in nodes])
all_cpus = os.sched_getaffinity(0)
print("affinity:", len(all_cpus), "cpus")


def parse(cl):
    s = set()
    for part in cl.split(","):
        if "-" in part:
            a, b = part.split("-"); s.update(range(int(a), int(b) + 1))
        elif part.strip():
            s.add(int(part))
    return s


N = 256 << 20
d = torch.empty(N, dtype=torch.uint8, device=dev)
for n in nodes:
    cpus = parse(open(n + "/cpulist").read().strip()) & all_cpus
    if not cpus:
        continue
    os.sched_setaffinity(0, cpus)
    h = torch.empty(N, dtype=torch.uint8).pin_memory()
    h.fill_(1)
    os.sched_setaffinity(0, all_cpus)
    for name, fn in (("H2D", lambda: d.copy_(h, non_blocking=True)), ("D2H", lambda: h.copy_(d, non_blocking=True))):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            fn()
        e1.record(); torch.cuda.synchronize()
        print(os.path.basename(n), name, f"{5 * N / (e0.elapsed_time(e1) * 1e-3) / 1e9:.1f} GB/s")
    del h

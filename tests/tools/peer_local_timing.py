"""Device time of the peer-exchange kernels when every destination is LOCAL memory: two simulated ranks on one GPU
(bench.py's N=2 slices: 10 M reads + 1 M variants per rank, 2 contigs), phases driven one by one, CUDA events around
each.  Compared with the laps of a real 2-GPU step (half the stores over NVLink) this separates what the kernels cost
from what the link costs.   python tests/tools/peer_local_timing.py [reads] [variants]"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from polars_bio_b200 import _native, dist as pbd  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
    m = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
    dev = torch.device("cuda:0")
    world, nc, T = 2, 2, 2
    rng = np.random.default_rng(1)
    L = 248_956_422

    def table(rows, length):
        c = torch.from_numpy(rng.integers(0, nc, rows).astype(np.int32)).to(dev)
        s = torch.from_numpy(rng.integers(0, L - length, rows).astype(np.int32)).to(dev)
        return c, s, s + length

    local = [[table(m, 1), table(n, 150)] for _ in range(world)]  # [rank][table]: indexed table first
    caps = [pbd._round_cap(int(1.3 * m) + 64), pbd._round_cap(int(1.3 * n) + 64)]
    arenas = [[torch.empty(16 * sum(caps), dtype=torch.uint8, device=dev) for _ in range(world)] for _ in range(2)]
    ctl = [torch.zeros(int(_native.lib().pbgpu_peer_ctl_bytes(world, T, nc)), dtype=torch.uint8, device=dev) for _ in range(world)]
    exs = [pbd.PeerExchange(nc, caps, dev, arenas=[[a.data_ptr() for a in arenas[p]] for p in range(2)],
                            ctls=[c.data_ptr() for c in ctl], world=world, rank=r) for r in range(world)]
    E = pbd.PeerExchange
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    laps = {}

    def timed(name, fn):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        laps.setdefault(name, []).append((a, b))

    for it in range(6):
        flush.fill_(it)
        for r in range(world):
            exs[r].begin_step(local[r])
        for r in range(world):
            timed("histograms+publish", lambda r=r: exs[r].phase(E.PH_HIST)) if r == 0 else exs[r].phase(E.PH_HIST)
        for r in range(world):
            timed("plan", lambda r=r: exs[r].phase(E.PH_PLAN)) if r == 0 else exs[r].phase(E.PH_PLAN)
        for r in range(world):
            for t in range(T):
                nm = f"table {t} ({'1M' if t == 0 else '10M'} rows): count+scan+scatter"
                timed(nm, lambda r=r, t=t: exs[r].phase(E.PH_SCATTER, t)) if r == 0 else exs[r].phase(E.PH_SCATTER, t)
                exs[r].phase(E.PH_SIGNAL, t)
        for r in range(world):
            for t in range(T):
                timed("wait (flags already raised)", lambda r=r, t=t: exs[r].phase(E.PH_WAIT, t)) if (r == 0 and t == 0) else exs[r].phase(E.PH_WAIT, t)
        torch.cuda.synchronize()
        for r in range(world):
            out, _ = exs[r].collect((exs[r].step - 1) & 1)
            assert out is not None and int(exs[r].status.item()) == 0
    res = {k: round(float(np.mean([a.elapsed_time(b) for a, b in v[2:]])), 4) for k, v in laps.items()}
    rows = n + m
    res["bytes_per_rank"] = 28 * rows
    res["grid_per_sm"] = os.environ.get("PBGPU_PEER_GRID", "2 (default)")
    print(json.dumps({"peer_local_timing_ms": res, "reads": n, "variants": m, "note": "rank 0 of 2 simulated ranks on one GPU, all stores local; L2 flushed per step"}))


if __name__ == "__main__":
    main()

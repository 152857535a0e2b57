"""BASELINE config 2 (the bench's `secondary` object) alone: one short run per environment (A/B of PBGPU_* knobs)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
r = bench.secondary_config2(torch.device("cuda:0"), steps=int(os.environ.get("PB_STEPS", "20")))
print(json.dumps({"env": {k: v for k, v in os.environ.items() if k.startswith("PBGPU_")}, **{k: (round(v, 5) if isinstance(v, float) else v) for k, v in r.items() if k.endswith("_ms") or k.startswith("ms_") or "frac" in k}}))

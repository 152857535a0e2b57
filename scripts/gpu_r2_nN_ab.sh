#!/bin/bash
# N-GPU A/B of exchange variants: bash scripts/gpu_r2_nN_ab.sh <tag> <N> "VAR=val ..." ...
TAG=$1; N=$2; shift 2
mkdir -p gpurun_out
if [ "$N" = "1" ]; then timeout 600 python -m pytest tests/test_gpu_peer.py -m gpu -x -q 2>&1 | tail -3; exit 0; fi
for v in "$@"; do
  n=$(echo "$v" | tr ' =,' '___')
  timeout 600 env $v python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --skip-e2e --skip-replicate > gpurun_out/${TAG}_${N}gpu_${n}.json 2> gpurun_out/${TAG}_${N}gpu_${n}.err; echo "[$v] rc=$?"
  python - gpurun_out/${TAG}_${N}gpu_${n}.json <<'PYEOF'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    c = d["config"]; r = d["roofline"]
    print("  N", d["n_gpus"], "ms/step %.3f" % d["ms_per_step"], "G pairs/s %.2f" % (d["value"] / 1e9), "exchange", c.get("exchange"), "xchg_ms %.3f" % c.get("exchange_ms_per_step", 0),
          "laps", {k[:28]: v for k, v in c.get("exchange_host_laps_ms", {}).items()}, "build %.3f" % r["index_build_ms"], "parity", d["parity_check"])
except Exception as e:
    print("  FAILED", e); print(open(sys.argv[1].replace(".json", ".err")).read()[-1500:])
PYEOF
done

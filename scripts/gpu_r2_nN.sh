#!/bin/bash
# N-GPU bench line(s): bash scripts/gpu_r2_nN.sh <tag> <N> [extra bench args]
TAG=$1; N=$2; shift 2
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 "$@" > gpurun_out/${TAG}_bench_${N}gpu.json 2> gpurun_out/${TAG}_bench_${N}gpu.err; echo rc=$?
python - gpurun_out/${TAG}_bench_${N}gpu.json <<'PYEOF'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
c = d["config"]; r = d["roofline"]
print("N", d["n_gpus"], "ms/step %.3f" % d["ms_per_step"], "G pairs/s %.2f" % (d["value"] / 1e9), "exchange", c.get("exchange"), "xchg_ms %.3f" % c.get("exchange_ms_per_step", 0),
      "laps", c.get("exchange_host_laps_ms"), "build %.3f" % r["index_build_ms"], {k.split("(")[0].strip(): round(v["ms"], 3) for k, v in r["all_stages"].items()},
      "parity", d["parity_check"], "e2e ms %.1f" % (d["e2e"]["ms_per_step"] if d.get("e2e") else -1), r.get("note"))
PYEOF
tail -4 gpurun_out/${TAG}_bench_${N}gpu.err

#!/bin/bash
# Peer-memory exchange round on N GPUs: simulated-rank kernel tests (1 GPU; both scatter grids), IPC + flags check vs
# the oracle, the bench line with the peer exchange and A/B variants (stream overlap, persistent scatter grid).
#   gpurun --gpus 2 --timeout 600 -- 'bash scripts/gpu_round_peer.sh r02a 2'
TAG=${1:-run}; N=${2:-2}
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== pytest tests/test_gpu_peer.py"
timeout 300 python -m pytest tests/test_gpu_peer.py -m gpu -x -q > $O/${TAG}_pytest_peer.log 2>&1
echo "pytest rc=$?"; tail -15 $O/${TAG}_pytest_peer.log
PBGPU_PEER_GRID=2 timeout 300 python -m pytest tests/test_gpu_peer.py -m gpu -x -q > $O/${TAG}_pytest_peer_grid2.log 2>&1
echo "pytest [PBGPU_PEER_GRID=2] rc=$?"; tail -3 $O/${TAG}_pytest_peer_grid2.log
echo "== dist_check ($N GPUs)"
timeout 300 $TR --master-port 29511 tests/tools/dist_check.py > $O/${TAG}_dist_check.log 2>&1
echo "dist_check rc=$?"; tail -12 $O/${TAG}_dist_check.log
echo "== bench --gpus $N (peer exchange, default settings)"
timeout 600 $TR --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 > $O/${TAG}_bench_${N}gpu.json 2> $O/${TAG}_bench_${N}gpu.err
echo "bench rc=$?"; tail -c 3200 $O/${TAG}_bench_${N}gpu.json; tail -5 $O/${TAG}_bench_${N}gpu.err
for v in ${VARIANTS:-PBGPU_BENCH_OVERLAP=1 PBGPU_BENCH_OVERLAP=1,PBGPU_PEER_GRID=2 PBGPU_PEER_GRID=2}; do
  n=$(echo "$v" | tr ',=' '__')
  echo "== bench --gpus $N [$v]"
  env $(echo $v | tr ',' ' ') timeout 600 $TR --master-port 29514 bench.py --gpus $N --steps 10 --warmup 3 --skip-e2e > $O/${TAG}_bench_${N}gpu_${n}.json 2> $O/${TAG}_bench_${N}gpu_${n}.err
  echo "bench rc=$?"; python - $O/${TAG}_bench_${N}gpu_${n}.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("ms/step %.4f" % d["ms_per_step"], "value %.3e" % d["value"], "exchange_ms %.4f" % d["config"]["exchange_ms_per_step"], d["config"]["exchange_host_laps_ms"])
except Exception as e:
    print("FAILED", e)
PY
  tail -3 $O/${TAG}_bench_${N}gpu_${n}.err
done
echo "== done"

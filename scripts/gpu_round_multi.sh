#!/bin/bash
# Multi-GPU round: sharded join vs the oracle, the bench line at N GPUs, the reference arm at N.
#   gpurun --gpus 2 --timeout 900 -- 'bash scripts/gpu_round_multi.sh r01z 2'
TAG=${1:-run}; N=${2:-2}
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== dist_check ($N GPUs)"
timeout 300 $TR --master-port 29511 tests/tools/dist_check.py > $O/${TAG}_dist_check.log 2>&1
echo "dist_check rc=$?"; tail -4 $O/${TAG}_dist_check.log
echo "== bench --gpus $N"
timeout 600 $TR --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 > $O/${TAG}_bench_${N}gpu.json 2> $O/${TAG}_bench_${N}gpu.err
echo "bench rc=$?"; tail -c 2500 $O/${TAG}_bench_${N}gpu.json; tail -5 $O/${TAG}_bench_${N}gpu.err
if [ -z "$SKIP_REF" ]; then
echo "== bench --impl reference --gpus $N"
timeout 600 $TR --master-port 29513 bench.py --impl reference --gpus $N --steps 1 --warmup 0 > $O/${TAG}_bench_ref_${N}gpu.json 2> $O/${TAG}_bench_ref_${N}gpu.err
echo "ref rc=$?"; tail -c 700 $O/${TAG}_bench_ref_${N}gpu.json
fi
echo "== done"

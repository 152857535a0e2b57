#!/bin/bash
# Peer-memory exchange round on N GPUs: simulated-rank kernel tests (1 GPU), IPC + NCCL check vs the oracle, the bench
# line with the peer exchange and, for A/B, with the NCCL all-to-all.
#   gpurun --gpus 2 --timeout 600 -- 'bash scripts/gpu_round_peer.sh r02a 2'
TAG=${1:-run}; N=${2:-2}
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== pytest tests/test_gpu_peer.py"
timeout 300 python -m pytest tests/test_gpu_peer.py -m gpu -x -q > $O/${TAG}_pytest_peer.log 2>&1
echo "pytest rc=$?"; tail -15 $O/${TAG}_pytest_peer.log
echo "== dist_check ($N GPUs)"
timeout 300 $TR --master-port 29511 tests/tools/dist_check.py > $O/${TAG}_dist_check.log 2>&1
echo "dist_check rc=$?"; tail -12 $O/${TAG}_dist_check.log
echo "== bench --gpus $N (peer exchange)"
timeout 600 $TR --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 > $O/${TAG}_bench_${N}gpu.json 2> $O/${TAG}_bench_${N}gpu.err
echo "bench rc=$?"; tail -c 2800 $O/${TAG}_bench_${N}gpu.json; tail -5 $O/${TAG}_bench_${N}gpu.err
echo "== bench --gpus $N (NCCL all-to-all, A/B)"
PBGPU_EXCHANGE=nccl timeout 600 $TR --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 --skip-e2e > $O/${TAG}_bench_${N}gpu_nccl.json 2> $O/${TAG}_bench_${N}gpu_nccl.err
echo "bench rc=$?"; tail -c 1500 $O/${TAG}_bench_${N}gpu_nccl.json; tail -5 $O/${TAG}_bench_${N}gpu_nccl.err
echo "== done"

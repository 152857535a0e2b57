#!/bin/bash
# Short single-GPU round: parity tests, one full bench line, build-stage trace of full-size config 3.
#   gpurun --timeout 900 -- 'bash scripts/gpu_round_b.sh r01x'
TAG=${1:-run}
O=gpurun_out
mkdir -p $O
echo "== pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -q ${PYTEST_ARGS:--x} > $O/${TAG}_pytest_gpu.log 2>&1
echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" $O/${TAG}_pytest_gpu.log | tail -40
grep -n -E "^E  " $O/${TAG}_pytest_gpu.log | head -60
echo "== bench"
timeout 600 python bench.py > $O/${TAG}_bench_1gpu.json 2> $O/${TAG}_bench_1gpu.err
echo "bench rc=$?"; tail -c 3500 $O/${TAG}_bench_1gpu.json; tail -5 $O/${TAG}_bench_1gpu.err
echo "== config 3 build trace"
PBGPU_TRACE_BUILD=1 PB_REPS=2 timeout 600 python tests/tools/scale_check.py 3 > $O/${TAG}_scale3.jsonl 2> $O/${TAG}_scale3_trace.err
echo "scale rc=$?"; cut -c1-900 $O/${TAG}_scale3.jsonl; grep -c "pbgpu build" $O/${TAG}_scale3_trace.err; tail -45 $O/${TAG}_scale3_trace.err
echo "== done"

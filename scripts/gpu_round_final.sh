#!/bin/bash
# Round-end style check on one GPU: smoke(), the whole GPU test-suite, the default bench line.
#   gpurun --timeout 200 -- 'bash scripts/gpu_round_final.sh r02e'
TAG=${1:-run}
O=gpurun_out
mkdir -p $O
echo "== smoke"
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" > $O/${TAG}_smoke.log 2>&1
echo "smoke rc=$?"; tail -2 $O/${TAG}_smoke.log
echo "== pytest -m gpu"
timeout 600 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest_gpu.log 2>&1
echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" $O/${TAG}_pytest_gpu.log | tail -10
grep -n -E "^E  " $O/${TAG}_pytest_gpu.log | head -30
echo "== bench"
timeout 400 python bench.py > $O/${TAG}_bench_1gpu.json 2> $O/${TAG}_bench_1gpu.err
echo "bench rc=$?"; tail -c 3500 $O/${TAG}_bench_1gpu.json; tail -3 $O/${TAG}_bench_1gpu.err
echo "== done"

#!/bin/bash
# round 2, call b: new scale tests + config-3 bench (N=1) with the round-1 kernels = the baseline of this round
O=gpurun_out; mkdir -p $O
echo "== pytest scale"; timeout 900 python -m pytest tests/test_gpu_scale.py -m gpu -x -q > $O/r2b_pytest_scale.log 2>&1; echo rc=$?; tail -15 $O/r2b_pytest_scale.log
echo "== bench N=1"; /usr/bin/time -v timeout 900 python bench.py --steps 10 --warmup 3 > $O/r2b_bench_1gpu.json 2> $O/r2b_bench_1gpu.err; echo rc=$?; tail -c 6000 $O/r2b_bench_1gpu.json; grep -E "Maximum resident|Elapsed" $O/r2b_bench_1gpu.err; tail -5 $O/r2b_bench_1gpu.err
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $O/r2b_bench_ref.json 2> $O/r2b_bench_ref.err; echo rc=$?; cat $O/r2b_bench_ref.json; tail -3 $O/r2b_bench_ref.err

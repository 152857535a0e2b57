#!/bin/bash
# One gpurun call: GPU parity tests, A/B of the env-selectable kernel variants, the bench line, the ncu launch list
# and one full capture of the provider / build kernels.  Everything lands in gpurun_out/<tag>_*.
#   gpurun --timeout 1500 -- 'bash scripts/gpu_round.sh r01n'
TAG=${1:-run}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.used --format=csv > $O/${TAG}_smi.txt 2>&1

echo "== pytest -m gpu"
timeout 1200 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest_gpu.log 2>&1
RC=$?
echo "pytest rc=$RC"; tail -5 $O/${TAG}_pytest_gpu.log
if [ $RC -ne 0 ]; then
  # pinpoint: which of the new kernels breaks parity?
  for v in "PBGPU_JDIR=search" "PBGPU_SORT=3k" "PBGPU_EMIT=walk" "PBGPU_JDIR=search PBGPU_SORT=3k PBGPU_EMIT=walk"; do
    n=$(echo "$v" | tr ' =' '__')
    timeout 600 env $v python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $O/${TAG}_pytest_${n}.log 2>&1
    echo "pytest [$v] rc=$?"; tail -3 $O/${TAG}_pytest_${n}.log
  done
fi

echo "== A/B variants (device-resident bench, no e2e)"
for v in ${VARIANTS:-PBGPU_X=default PBGPU_JDIR=search PBGPU_SORT=3k PBGPU_EMIT=walk PBGPU_SYNC=memcpy}; do
  n=$(echo "$v" | tr ' =' '__')
  timeout 300 env $v python bench.py --steps 20 --warmup 3 --no-cpu-baseline --skip-e2e > $O/${TAG}_ab_${n}.json 2> $O/${TAG}_ab_${n}.err
  python - "$O/${TAG}_ab_${n}.json" "$v" <<'EOF'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = d["roofline"]
    print(sys.argv[2], "ms/step %.4f" % d["ms_per_step"], "build %.4f" % r["index_build_ms"],
          {k.split()[0]: round(v["ms"], 4) for k, v in r["all_kernels"].items()}, r["step_stage_ms"])
except Exception as e:
    print(sys.argv[2], "FAILED", e)
EOF
done

if [ -z "$SKIP_BENCH" ]; then
echo "== bench (full line)"
timeout 900 python bench.py > $O/${TAG}_bench_1gpu.json 2> $O/${TAG}_bench_1gpu.err
tail -c 3000 $O/${TAG}_bench_1gpu.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/${TAG}_bench_ref.json 2> $O/${TAG}_bench_ref.err
tail -c 600 $O/${TAG}_bench_ref.json
fi

if [ -z "$SKIP_NCU_LIST" ]; then
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --skip-e2e > $O/${TAG}_ncu_launches.log 2>&1
echo "ncu launches rc=$?"
fi
if [ -z "$SKIP_NCU" ]; then
echo "== ncu full capture"
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:'overlap_emit|count_overlaps_fast|overlap_count_fast|rs_onesweep|rs_hist_all|jdir_|unpack_sorted|make_start_keys|build_stats' \
    -s 13 -c 14 -f -o $O/${TAG}_prof python bench.py --steps 2 --warmup 3 --no-cpu-baseline --skip-e2e > $O/${TAG}_ncu_full.log 2>&1
echo "ncu full rc=$?"; ls -la $O/${TAG}_prof.ncu-rep
fi

if [ -n "$SCALE_CONFIGS" ]; then
echo "== full-size configs $SCALE_CONFIGS"
timeout 900 python tests/tools/scale_check.py $SCALE_CONFIGS > $O/${TAG}_scale.jsonl 2> $O/${TAG}_scale.err
cat $O/${TAG}_scale.jsonl | cut -c1-1200
fi
if [ -n "$SCALE_NCU" ]; then
echo "== ncu launch list of full-size config $SCALE_NCU"
PB_REPS=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/${TAG}_scale${SCALE_NCU}_launches.csv \
    python tests/tools/scale_check.py $SCALE_NCU > $O/${TAG}_scale${SCALE_NCU}_ncu.log 2>&1
echo "rc=$?"
fi
if [ -n "$E2E_TRACE" ]; then
echo "== e2e trace"
timeout 300 python scripts/e2e_trace.py > $O/${TAG}_e2e_trace.log 2>&1
tail -40 $O/${TAG}_e2e_trace.log
fi
echo "== done"

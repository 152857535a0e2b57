#!/bin/bash
# round 2: one gpurun call = [pytest -m gpu] + [bench N=1] + optional extras, everything under gpurun_out/<tag>_*
#   gpurun --timeout 2400 -- 'bash scripts/gpu_r2.sh r2c pytest bench trace'
TAG=${1:-r2}; shift
O=gpurun_out; mkdir -p $O
for what in "$@"; do
case $what in
pytest)
  echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest_gpu.log 2>&1; echo rc=$?; tail -12 $O/${TAG}_pytest_gpu.log;;
pytest_scale)
  echo "== pytest scale"; timeout 900 python -m pytest tests/test_gpu_scale.py -m gpu -x -q > $O/${TAG}_pytest_scale.log 2>&1; echo rc=$?; tail -12 $O/${TAG}_pytest_scale.log;;
ids)
  echo "== ids check"; timeout 600 python tests/tools/ids_check.py > $O/${TAG}_ids_check.log 2>&1; echo rc=$?; tail -6 $O/${TAG}_ids_check.log; PBGPU_BIN=1 timeout 600 python tests/tools/ids_check.py > $O/${TAG}_ids_check_bins.log 2>&1; echo rc=$?; tail -3 $O/${TAG}_ids_check_bins.log;;
payload)
  echo "== pytest payload"; timeout 900 python -m pytest tests/test_gpu_payload.py tests/test_gpu_api.py -m gpu -x -q > $O/${TAG}_pytest_payload.log 2>&1; echo rc=$?; tail -25 $O/${TAG}_pytest_payload.log;;
stream)
  echo "== pytest stream"; timeout 900 python -m pytest tests/test_gpu_stream.py tests/test_gpu_api.py -m gpu -x -q > $O/${TAG}_pytest_stream.log 2>&1; echo rc=$?; tail -25 $O/${TAG}_pytest_stream.log;;
bins)
  echo "== bins check"; PBGPU_BIN=1 timeout 900 python tests/tools/bins_check.py > $O/${TAG}_bins_check.log 2>&1; echo rc=$?; tail -8 $O/${TAG}_bins_check.log;;
pytest_parity)
  echo "== pytest parity"; timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py tests/test_gpu_unary.py -m gpu -x -q > $O/${TAG}_pytest_parity.log 2>&1; echo rc=$?; tail -12 $O/${TAG}_pytest_parity.log;;
bench)
  echo "== bench N=1"; timeout 1200 python bench.py --steps 10 --warmup 3 > $O/${TAG}_bench_1gpu.json 2> $O/${TAG}_bench_1gpu.err; echo rc=$?; tail -c 7000 $O/${TAG}_bench_1gpu.json; tail -5 $O/${TAG}_bench_1gpu.err;;
bench_dev)
  echo "== bench N=1 (device only)"; timeout 600 python bench.py --steps 10 --warmup 3 --skip-e2e --no-cpu-baseline --skip-secondary > $O/${TAG}_bench_dev.json 2> $O/${TAG}_bench_dev.err; echo rc=$?; tail -c 5000 $O/${TAG}_bench_dev.json; tail -5 $O/${TAG}_bench_dev.err;;
bench_dev2)
  echo "== bench N=1 (device only, with the config-2 secondary and the parity check)"; timeout 900 python bench.py --steps 10 --warmup 3 --skip-e2e --no-cpu-baseline > $O/${TAG}_bench_dev2.json 2> $O/${TAG}_bench_dev2.err; echo rc=$?; python scripts/bench_brief.py $O/${TAG}_bench_dev2.json; tail -5 $O/${TAG}_bench_dev2.err;;
ab2)
  echo "== A/B variants (device-resident bench with the config-2 secondary, no e2e)"
  for v in ${AB_VARIANTS:-PBGPU_X=default}; do
    n=$(echo "$v" | tr ' =,/' '____'); vv=$(echo "$v" | tr ',' ' ')
    timeout 600 env $vv python bench.py --steps 8 --warmup 3 --skip-e2e --no-cpu-baseline --skip-parity > $O/${TAG}_ab_${n}.json 2> $O/${TAG}_ab_${n}.err
    echo "-- $v"; python scripts/bench_brief.py $O/${TAG}_ab_${n}.json || tail -5 $O/${TAG}_ab_${n}.err
  done;;
e2etrace)
  echo "== e2e trace (config 3 through the public API)"; timeout 600 python scripts/e2e_trace3.py > /dev/null 2> $O/${TAG}_e2e_trace3.txt; echo rc=$?; grep -v "^\[pbgpu\]   \|get_next" $O/${TAG}_e2e_trace3.txt | tail -16;;
share8)
  PB_WORLD=8 PB_RANK=0 timeout 300 python tests/tools/rank_share.py > $O/${TAG}_rank_share_w8.json 2> $O/${TAG}_rank_share_w8.err; echo rc=$?; cat $O/${TAG}_rank_share_w8.json; tail -3 $O/${TAG}_rank_share_w8.err;;
ab)
  echo "== A/B variants (device-resident bench, no e2e)"
  for v in ${AB_VARIANTS:-PBGPU_X=default}; do
    n=$(echo "$v" | tr ' =,/' '____'); vv=$(echo "$v" | tr ',' ' ')
    timeout 600 env $vv python bench.py --steps 8 --warmup 3 --skip-e2e --no-cpu-baseline --skip-secondary --skip-parity > $O/${TAG}_ab_${n}.json 2> $O/${TAG}_ab_${n}.err
    python - "$O/${TAG}_ab_${n}.json" "$v" <<'PYEOF'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = d["roofline"]
    print(sys.argv[2], "ms/step %.3f" % d["ms_per_step"], "build %.3f" % r["index_build_ms"],
          {k.split("(")[0].strip(): round(v["ms"], 3) for k, v in r["all_stages"].items()}, "bin %.3f unbin %.3f" % (r.get("bin_ms", 0), r.get("unbin_ms", 0)))
except Exception as e:
    print(sys.argv[2], "FAILED", e, open(sys.argv[1].replace(".json", ".err")).read()[-600:])
PYEOF
  done;;
generic)
  echo "== parity with the generic build front half"; PBGPU_BUILD=generic timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py -m gpu -x -q > $O/${TAG}_pytest_generic.log 2>&1; echo rc=$?; tail -4 $O/${TAG}_pytest_generic.log;;
trace)
  echo "== build trace (config 3)"; PBGPU_TRACE_BUILD=1 PB_REPS=2 timeout 600 python tests/tools/scale_check.py 3 > $O/${TAG}_scale3.jsonl 2> $O/${TAG}_build_trace.txt; echo rc=$?; tail -14 $O/${TAG}_build_trace.txt; cat $O/${TAG}_scale3.jsonl;;
ref)
  echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $O/${TAG}_bench_ref.json 2> $O/${TAG}_bench_ref.err; echo rc=$?; cat $O/${TAG}_bench_ref.json;;
launches)
  echo "== ncu launch list (config 3: one build + count_overlaps + overlap, cold)"
  PB_REPS=2 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_config3_launches.csv python tests/tools/profile_config3.py > $O/${TAG}_launches.log 2>&1; echo rc=$?; tail -3 $O/${TAG}_launches.log;;
prof)
  echo "== ncu --set full (config 3, one launch of every hot kernel)"
  timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"${PROF_REGEX:-prep_kernel|rs_onesweep|unpack_sorted|pmax_lookback|jdir_|bin_hist|bin_partition|binned_|unbin_|emit_staged|count_overlaps_fast|overlap_count_fast|overlap_emit}" -c ${PROF_COUNT:-24} -o $O/${TAG}_config3_prof -f python tests/tools/profile_config3.py > $O/${TAG}_prof.log 2>&1; echo rc=$?; tail -3 $O/${TAG}_prof.log; ls -la $O/${TAG}_config3_prof.ncu-rep;;
esac
done

#!/bin/bash
# one rank's share of config 3 at world size W on one GPU: timings, build trace, ncu launch list
TAG=${1:-r2v}; W=${2:-8}
O=gpurun_out; mkdir -p $O
PB_WORLD=$W PB_RANK=0 timeout 300 python tests/tools/rank_share.py > $O/${TAG}_rank_share_w$W.json 2> $O/${TAG}_rank_share_w$W.err; echo rc=$?; cat $O/${TAG}_rank_share_w$W.json; tail -3 $O/${TAG}_rank_share_w$W.err
PB_WORLD=$W PB_RANK=0 PB_REPS=3 PBGPU_TRACE_BUILD=1 timeout 300 python tests/tools/rank_share.py > /dev/null 2> $O/${TAG}_rank_share_w${W}_trace.txt; tail -12 $O/${TAG}_rank_share_w${W}_trace.txt
PB_WORLD=$W PB_RANK=0 PB_REPS=2 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_rank_share_w${W}_launches.csv python tests/tools/rank_share.py > $O/${TAG}_launches.log 2>&1; echo rc=$?
python scripts/launch_shares.py $O/${TAG}_rank_share_w${W}_launches.csv | head -45

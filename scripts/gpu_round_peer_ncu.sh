#!/bin/bash
# ncu on the peer-exchange kernels from ONE process (simulated ranks on one GPU; never profile a multi-rank command):
# launch list + one full capture of the scatter / count / histogram kernels.
#   gpurun --timeout 600 -- 'bash scripts/gpu_round_peer_ncu.sh r03a'
TAG=${1:-run}; O=gpurun_out; mkdir -p $O
timeout 200 python tests/tools/peer_local_timing.py > $O/${TAG}_peer_local.json 2> $O/${TAG}_peer_local.err
echo "timing rc=$?"; cat $O/${TAG}_peer_local.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/${TAG}_peer_launches.csv \
    python tests/tools/peer_local_timing.py > $O/${TAG}_peer_ncu_list.log 2>&1
echo "ncu list rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'peer_scatter|peer_block_count|peer_hist_publish|peer_plan' \
    -s 12 -c 12 -f -o $O/${TAG}_peer_prof python tests/tools/peer_local_timing.py > $O/${TAG}_peer_ncu_full.log 2>&1
echo "ncu full rc=$?"; ls -la $O/${TAG}_peer_prof.ncu-rep
echo "== done"

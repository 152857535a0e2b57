#!/bin/bash
# Local (single-GPU, simulated ranks) device time of the peer-exchange kernels + their ncu launch list.
TAG=${1:-run}; O=gpurun_out; mkdir -p $O
timeout 100 python tests/tools/peer_local_timing.py > $O/${TAG}_peer_local.json 2> $O/${TAG}_peer_local.err
echo "rc=$?"; cat $O/${TAG}_peer_local.json; tail -3 $O/${TAG}_peer_local.err
PBGPU_PEER_GRID=0 timeout 100 python tests/tools/peer_local_timing.py > $O/${TAG}_peer_local_grid0.json 2>> $O/${TAG}_peer_local.err
echo "rc=$?"; cat $O/${TAG}_peer_local_grid0.json
echo "== done"

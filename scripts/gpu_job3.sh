#!/bin/bash
TAG=${1:-r1c}
O=gpurun_out/$TAG
mkdir -p $O
timeout 300 python scripts/march_probe.py > $O/march_probe.log 2>&1
timeout 200 python scripts/pair_phases.py > $O/pair_phases_74.log 2>&1
RNERF_PAIR_LIMIT=16 timeout 200 python scripts/pair_phases.py > $O/pair_phases_16.log 2>&1
cat $O/march_probe.log; grep -A24 "layer pair" $O/pair_phases_74.log | cut -c1-80; grep -A24 "layer pair" $O/pair_phases_16.log | cut -c1-80

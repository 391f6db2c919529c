#!/bin/bash
TAG=${1:-r1e}
O=gpurun_out/$TAG
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
timeout 300 python scripts/mlp_dbg_probe.py 0 7 > $O/mlp_dbg.log 2>&1
timeout 200 python scripts/pair_phases.py > $O/pair_phases.log 2>&1
tail -4 $O/pytest_gpu.log; cat $O/mlp_dbg.log; grep -A24 "layer pair" $O/pair_phases.log | cut -c1-80

#!/bin/bash
TAG=${1:-r1f}
O=gpurun_out/$TAG
mkdir -p $O
timeout 600 python scripts/mlp_dbg_probe.py 0 7 8 16 24 32 39 64 103 0 > $O/mlp_dbg.log 2>&1
RNERF_MLP_KERNEL=single timeout 200 python scripts/mlp_dbg_probe.py 0 > $O/mlp_single.log 2>&1
cat $O/mlp_dbg.log $O/mlp_single.log

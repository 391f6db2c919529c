#!/bin/bash
TAG=${1:-r1g}
O=gpurun_out/$TAG
mkdir -p $O
for V in base pipelined_ld base pipelined_ld; do
  echo "== $V" >> $O/mlp_ab.log
  RNERF_LIB=$PWD/samplenerfro_b200/build/variants/$V.so timeout 200 python scripts/mlp_dbg_probe.py 0 >> $O/mlp_ab.log 2>&1
done
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
cat $O/mlp_ab.log; tail -3 $O/pytest_gpu.log

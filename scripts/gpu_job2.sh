#!/bin/bash
TAG=${1:-r1b}
O=gpurun_out/$TAG
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
timeout 300 python scripts/hbm_probe.py > $O/hbm.log 2>&1
timeout 300 python scripts/perf_probe.py > $O/probe.log 2>&1
timeout 300 python scripts/pair_phases.py > $O/pair_phases.log 2>&1
for K in march_kernel resample_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -c 1 --launch-skip 12 -f -o $O/$K \
     python scripts/perf_probe.py > $O/ncu_$K.log 2>&1
done
tail -5 $O/pytest_gpu.log; cat $O/hbm.log $O/probe.log; tail -30 $O/pair_phases.log

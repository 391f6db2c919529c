#!/bin/bash
# Visit: full GPU parity suite, stage probe (composite fast math), all-stage march timing, bench.
TAG=${1:-r1i}
O=gpurun_out/$TAG
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -25 $O/pytest_gpu.log
timeout 300 python scripts/perf_probe.py > $O/probe.log 2>&1; cat $O/probe.log
timeout 300 python scripts/all_stage_probe.py > $O/all_stage.log 2>&1; cat $O/all_stage.log

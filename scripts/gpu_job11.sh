#!/bin/bash
TAG=${1:-r1n}
O=gpurun_out/$TAG
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_golden.py tests/test_gpu_model.py -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -8 $O/pytest_gpu.log
timeout 300 python scripts/all_stage_probe.py > $O/all_stage.log 2>&1; cat $O/all_stage.log

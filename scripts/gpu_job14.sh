#!/bin/bash
TAG=${1:-r1q}
O=gpurun_out/$TAG
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_fullsize.py -x -q > $O/pytest_full.log 2>&1; echo "pytest rc=$?" >> $O/pytest_full.log
tail -30 $O/pytest_full.log

#!/bin/bash
# Run selected GPU test files (and optionally more commands) on the box:  bash scripts/gpu_tests.sh <tag> <pytest args...>
TAG=$1; shift
O=gpurun_out/$TAG
mkdir -p $O
timeout 900 python -m pytest "$@" -q -s --durations=6 > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -60 $O/pytest.log

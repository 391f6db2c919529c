#!/bin/bash
TAG=${1:-r1d}
O=gpurun_out/$TAG
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q -k "march or resample or select or model or golden" > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
timeout 300 python scripts/march_probe.py > $O/march_probe.log 2>&1
timeout 300 python scripts/mlp_dbg_probe.py > $O/mlp_dbg.log 2>&1
timeout 300 python bench.py --no-cpu-baseline --chunk 640000 > $O/bench_chunk640k.json 2>$O/bench.err
tail -4 $O/pytest_gpu.log; cat $O/march_probe.log $O/mlp_dbg.log; cat $O/bench_chunk640k.json

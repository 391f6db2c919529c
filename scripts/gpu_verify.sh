#!/bin/bash
# HEAD verification on one B200: full GPU parity suite, smoke(), optionally one default bench line.
#   bash scripts/gpu_verify.sh <tag> [bench]
TAG=${1:-r1q}
O=gpurun_out/$TAG
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q --durations=8 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -16 $O/pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
if [ "$2" = "bench" ]; then
  timeout 400 python bench.py > $O/bench.json 2> $O/bench.err; tail -2 $O/bench.err; cat $O/bench.json
fi

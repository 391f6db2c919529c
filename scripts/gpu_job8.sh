#!/bin/bash
# Visit: full GPU parity suite, all-stage probe, ball config, march sweep.
TAG=${1:-r1j}
O=gpurun_out/$TAG
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -25 $O/pytest_gpu.log
timeout 300 python scripts/all_stage_probe.py > $O/all_stage.log 2>&1; cat $O/all_stage.log
timeout 300 python scripts/config_bench.py --config ball > $O/ball.json 2> $O/ball.err; tail -3 $O/ball.err; cat $O/ball.json
timeout 600 python scripts/config_bench.py --config sweep > $O/sweep.jsonl 2> $O/sweep.err; tail -3 $O/sweep.err; tail -4 $O/sweep.jsonl

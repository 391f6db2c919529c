#!/bin/bash
# Verification visit: parity tests, smoke, per-stage probe, headline bench (both arms), training bench, launch list.
TAG=${1:-r1c}
O=gpurun_out/$TAG
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log
timeout 300 python scripts/perf_probe.py > $O/probe.log 2>&1
timeout 600 python bench.py > $O/bench.json 2> $O/bench.err
timeout 300 python scripts/train_bench.py > $O/train.json 2> $O/train.err
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
   --log-file $O/launches_bench.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $O/ncu_bench.log 2>&1
tail -5 $O/pytest_gpu.log; tail -2 $O/smoke.log; cat $O/probe.log; cat $O/bench.json; cat $O/train.json; cat $O/bench_ref.json

#!/bin/bash
# HEAD verification on one B200: full GPU parity suite, smoke(), the default bench line (both arms), all-stage numbers.
TAG=${1:-r4g}
O=gpurun_out/$TAG
mkdir -p $O
timeout 1100 python -m pytest tests -m gpu -q --durations=6 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log; tail -4 $O/pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > $O/smoke.log 2>&1; tail -1 $O/smoke.log
timeout 500 python bench.py --steps 10 --warmup 3 > $O/bench.json 2> $O/bench.err; tail -2 $O/bench.err; cut -c1-400 $O/bench.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2>> $O/bench.err; cut -c1-200 $O/bench_reference.json
timeout 300 python scripts/train_bench.py --stage all --steps 10 --warmup 5 > $O/train_all.json 2>> $O/bench.err; cut -c1-200 $O/train_all.json
timeout 300 python scripts/all_stage_probe.py > $O/all_stage_probe.txt 2>&1; tail -3 $O/all_stage_probe.txt | cut -c1-300

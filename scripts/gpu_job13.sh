#!/bin/bash
TAG=${1:-r1p}
O=gpurun_out/$TAG
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_train.py -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -4 $O/pytest_gpu.log
timeout 300 python scripts/train_bench.py --steps 20 --warmup 5 > $O/train.json 2> $O/train.err; tail -2 $O/train.err; cat $O/train.json
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
   --log-file $O/launches_train.csv python scripts/train_bench.py --steps 1 --warmup 4 > $O/ncu_train.log 2>&1
python scripts/launch_summary.py $O/launches_train.csv > $O/launches_train_summary.txt; head -8 $O/launches_train_summary.txt

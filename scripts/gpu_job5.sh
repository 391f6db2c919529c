#!/bin/bash
# Training-path visit: train tests, training bench (graph + eager), launch list of one replayed step.
TAG=${1:-r1e}
O=gpurun_out/$TAG
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_train.log 2>&1; echo "pytest rc=$?" >> $O/pytest_train.log
tail -30 $O/pytest_train.log
timeout 300 python scripts/train_bench.py --steps 20 --warmup 5 > $O/train.json 2> $O/train.err; tail -3 $O/train.err
cat $O/train.json
timeout 300 python scripts/train_bench.py --steps 20 --warmup 5 --eager > $O/train_eager.json 2> $O/train_eager.err
cat $O/train_eager.json
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
   --log-file $O/launches_train.csv python scripts/train_bench.py --steps 1 --warmup 4 > $O/ncu_train.log 2>&1
python scripts/launch_summary.py $O/launches_train.csv | head -14

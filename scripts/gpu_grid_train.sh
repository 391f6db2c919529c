#!/bin/bash
# learned-grid extension: tests, training bench line at G=512 (radiance stage), launch list of one replayed step.
TAG=${1:-r2g}
O=gpurun_out/$TAG
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_all_stage_train.py -x -q -k "grid" > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -5 $O/pytest.log
timeout 300 python scripts/train_bench.py --learn-grid --steps 10 --warmup 5 > $O/train_grid.json 2> $O/train_grid.err; tail -3 $O/train_grid.err; cat $O/train_grid.json
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
   --log-file $O/launches_train_grid.csv python scripts/train_bench.py --learn-grid --steps 1 --warmup 4 > $O/ncu_train_grid.log 2>&1
python scripts/launch_summary.py $O/launches_train_grid.csv > $O/launches_train_grid_summary.txt; head -12 $O/launches_train_grid_summary.txt

#!/bin/bash
# "all"-stage: parity tests, training bench line, launch list of one replayed step.
TAG=${1:-r1t}
O=gpurun_out/$TAG
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_all_stage_train.py tests/test_gpu_golden.py tests/test_gpu_model.py -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -5 $O/pytest.log
timeout 300 python scripts/train_bench.py --stage all --steps 10 --warmup 5 > $O/train_all.json 2> $O/train_all.err; tail -3 $O/train_all.err; cat $O/train_all.json
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
   --log-file $O/launches_train_all.csv python scripts/train_bench.py --stage all --steps 1 --warmup 4 > $O/ncu_train_all.log 2>&1
python scripts/launch_summary.py $O/launches_train_all.csv > $O/launches_train_all_summary.txt; head -6 $O/launches_train_all_summary.txt

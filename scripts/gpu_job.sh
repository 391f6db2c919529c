#!/bin/bash
# One GPU-box visit: parity tests, per-stage probe, headline bench, training bench, ncu launch list + full captures.
# Usage (from the repo root on the box): bash scripts/gpu_job.sh [tag]
TAG=${1:-r1}
O=gpurun_out/$TAG
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
timeout 300 python scripts/perf_probe.py > $O/probe.log 2>&1
timeout 600 python bench.py > $O/bench.json 2> $O/bench.err
timeout 300 python scripts/train_bench.py > $O/train.json 2> $O/train.err
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
   --log-file $O/launches_bench.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $O/ncu_bench.log 2>&1
for K in encmlp_pair_kernel march_kernel composite_fwd_kernel resample_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -c 1 --launch-skip 2 -f -o $O/$K \
     python scripts/perf_probe.py --rays 65536 > $O/ncu_$K.log 2>&1
done
tail -3 $O/pytest_gpu.log; cat $O/probe.log; cat $O/bench.json; cat $O/train.json

#!/bin/bash
TAG=${1:-r1k}
O=gpurun_out/$TAG
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_rays_ckpt.py -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -5 $O/pytest_gpu.log
timeout 300 python scripts/all_stage_probe.py > $O/all_stage.log 2>&1; cat $O/all_stage.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:march_kernel -c 1 --launch-skip 1 -f -o $O/march_all \
     python scripts/all_stage_probe.py --rays 65536 > $O/ncu_march_all.log 2>&1; tail -3 $O/ncu_march_all.log

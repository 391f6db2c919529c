#!/bin/bash
# Visit: all-stage parity + probe after the compaction rewrite; ncu --set full captures of the kernels changed this round.
TAG=${1:-r1m}
O=gpurun_out/$TAG
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_golden.py tests/test_gpu_model.py -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -5 $O/pytest_gpu.log
timeout 300 python scripts/all_stage_probe.py > $O/all_stage.log 2>&1; cat $O/all_stage.log
NCU="ncu --set full --clock-control none --import-source on -c 1 -f"
timeout 600 $NCU -k regex:march_kernel --launch-skip 8 -o $O/march_all_kernel python scripts/all_stage_probe.py --rays 65536 > $O/ncu_march_all.log 2>&1
timeout 600 $NCU -k regex:composite_fwd_kernel --launch-skip 12 -o $O/composite_fwd_kernel python scripts/perf_probe.py --rays 65536 > $O/ncu_composite.log 2>&1
timeout 600 $NCU -k regex:mlp_wgrad_kernel --launch-skip 49 -o $O/mlp_wgrad_kernel python scripts/train_bench.py --steps 1 --warmup 2 --eager > $O/ncu_wgrad.log 2>&1
timeout 600 $NCU -k regex:mlp_dgrad_kernel --launch-skip 4 -o $O/mlp_dgrad_kernel python scripts/train_bench.py --steps 1 --warmup 2 --eager > $O/ncu_dgrad.log 2>&1
timeout 600 $NCU -k regex:encmlp_kernel --launch-skip 5 -o $O/encmlp_train_kernel python scripts/train_bench.py --steps 1 --warmup 2 --eager > $O/ncu_fwdtrain.log 2>&1
ls -la $O/*.ncu-rep

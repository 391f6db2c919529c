#!/bin/bash
# 2-GPU visit: distributed training check (incl. graph-captured NCCL), training bench and headline bench at N=2.
TAG=${1:-r1h}
O=gpurun_out/$TAG
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR scripts/train_2gpu_check.py > $O/train_2gpu_check.log 2>&1; echo "rc=$?" >> $O/train_2gpu_check.log
tail -4 $O/train_2gpu_check.log
timeout 300 $TR scripts/train_bench.py --steps 20 --warmup 5 > $O/train_n2.json 2> $O/train_n2.err; tail -2 $O/train_n2.err; cat $O/train_n2.json
timeout 600 $TR bench.py --gpus 2 --steps 3 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err; tail -2 $O/bench_n2.err; cat $O/bench_n2.json

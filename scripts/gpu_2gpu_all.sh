#!/bin/bash
# 2 x B200: "all"-stage gradient mean vs single-GPU full batch, graph replay with in-graph NCCL, N=2 training bench.
TAG=${1:-r2b}
O=gpurun_out/$TAG
mkdir -p $O
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 \
   scripts/train_2gpu_check.py all > $O/train_2gpu_check_all.txt 2>&1; echo "rc=$?" >> $O/train_2gpu_check_all.txt; grep -v Warning $O/train_2gpu_check_all.txt | tail -5
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 \
   scripts/train_bench.py --stage all --steps 10 --warmup 5 > $O/train_all_n2.json 2> $O/train_all_n2.err; tail -2 $O/train_all_n2.err; cat $O/train_all_n2.json

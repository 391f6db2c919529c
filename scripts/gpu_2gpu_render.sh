#!/bin/bash
# 2 x B200: ball.gin-shaped frame row-sharded over the ranks with the band all-gather, and on one GPU.
TAG=${1:-r2d}
O=gpurun_out/$TAG
mkdir -p $O
timeout 300 python scripts/config_bench.py --config ball > $O/ball_n1.json 2> $O/ball_n1.err; tail -2 $O/ball_n1.err; cat $O/ball_n1.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29621 \
   scripts/config_bench.py --config ball > $O/ball_n2.json 2> $O/ball_n2.err; grep -v "Warning\|OMP\|\*\*\*" $O/ball_n2.err | tail -3; cat $O/ball_n2.json

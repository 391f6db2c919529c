#!/bin/bash
# Evidence pass at HEAD on one B200 (reduced here by scripts/ncu_multi_summary.py / launch_summary.py):
#   render: launch list + `ncu --set full` of every kernel of the frame at the bench's own launch sizes (640 000 rays)
#   all-stage: `--set full` of march_tc_kernel (band of 163 840 rays) and of the reverse sweep (4096 random pixels)
#   training: launch list + `--set full` of the forward / dgrad / wgrad / head kernels of one eager 4096-ray step
#   bash scripts/gpu_evidence.sh <tag>
TAG=${1:-r4a}
O=gpurun_out/$TAG
mkdir -p $O
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv \
   --log-file $O/launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-train > $O/launches_bench.log 2>&1
timeout 1200 ncu --set full --import-source on --clock-control none --profile-from-start off \
   -k regex:'march_kernel|resample_kernel|encmlp_pair_kernel|composite_fwd_kernel|select_kernel|bkgd_mlp_kernel' -c 8 -f -o $O/render \
   python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-train > $O/ncu_render.log 2>&1
ncu -i $O/render.ncu-rep --page raw --csv > $O/render_raw.csv
timeout 900 ncu --set full --import-source on --clock-control none -k regex:'march_tc_kernel' -c 1 -f -o $O/march_tc \
   python scripts/all_stage_ncu.py --rays 163840 > $O/ncu_march_tc.log 2>&1
ncu -i $O/march_tc.ncu-rep --page raw --csv > $O/march_tc_raw.csv
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
   --log-file $O/launches_train.csv python scripts/train_bench.py --steps 1 --warmup 4 > $O/ncu_train_list.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none --profile-from-start off \
   -k regex:'mlp_dgrad_kernel|mlp_wgrad_kernel|encmlp_pair_kernel|mlp_head_grad|composite_bwd' -c 10 -f -o $O/train \
   python scripts/train_bench.py --steps 1 --warmup 4 --eager > $O/ncu_train.log 2>&1
ncu -i $O/train.ncu-rep --page raw --csv > $O/train_raw.csv
rm -f $O/*.ncu-rep
ls -la $O

#!/bin/bash
# Evidence pass on one B200 at the bench's own launch sizes (640 000-ray frame): launch list of the timed region,
# then ONE `ncu --set full` report holding the march / resample / enc+MLP (coarse, fine) / composite (coarse, fine) launches.
#   bash scripts/gpu_ncu_render.sh <tag>        (reduced here by scripts/ncu_multi_summary.py)
TAG=${1:-r3a}
O=gpurun_out/$TAG
mkdir -p $O
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv \
   --log-file $O/launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/launches_bench.log 2>&1
timeout 1200 ncu --set full --import-source on --clock-control none --profile-from-start off \
   -k regex:'march_kernel|resample_kernel|encmlp_pair_kernel|composite_fwd_kernel|select_kernel|bkgd_mlp_kernel' -c 8 -f -o $O/render \
   python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $O/ncu_render.log 2>&1
tail -3 $O/ncu_render.log
ncu -i $O/render.ncu-rep --page raw --csv > $O/render_raw.csv
ls -la $O

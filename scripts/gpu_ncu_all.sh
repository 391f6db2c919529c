#!/bin/bash
# ncu --set full of the "all"-stage forward march (band of image rows) and its reverse sweep (random training batch);
# the reports are reduced to csv / per-line text on the box (they are too big to bring back).
TAG=${1:-r1w}
O=gpurun_out/$TAG
mkdir -p $O /tmp/ncu
timeout 600 ncu --set full --import-source on --clock-control none -k regex:'march_kernel' -f -o /tmp/ncu/fwd \
   python scripts/all_stage_ncu.py --rays 65536 > $O/ncu_fwd.log 2>&1; tail -1 $O/ncu_fwd.log
ncu -i /tmp/ncu/fwd.ncu-rep --page raw --csv > $O/march_all_fwd_raw.csv
python scripts/ncu_lines.py /tmp/ncu/fwd.ncu-rep samplenerfro_b200/build/march.o march_kernelILi2ELb1ELb1 > $O/march_all_fwd_lines.txt 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:'march_all_bwd' -f -o /tmp/ncu/bwd \
   python scripts/all_stage_ncu.py --rays 4096 --random > $O/ncu_bwd.log 2>&1; tail -1 $O/ncu_bwd.log
ncu -i /tmp/ncu/bwd.ncu-rep --page raw --csv > $O/march_all_bwd_raw.csv
python scripts/ncu_lines.py /tmp/ncu/bwd.ncu-rep samplenerfro_b200/build/march_bwd.o march_all_bwd_kernelILb1 > $O/march_all_bwd_lines.txt 2>&1
ls -la $O

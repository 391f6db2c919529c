#!/bin/bash
# Training-step breakdown: launch list of one timed train step (all kernels, ours + torch's).
TAG=${1:-r1d}
O=gpurun_out/$TAG
mkdir -p $O
timeout 300 python scripts/train_bench.py > $O/train.json 2> $O/train.err
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
   --log-file $O/launches_train.csv python scripts/train_bench.py --steps 1 --warmup 2 > $O/ncu_train.log 2>&1
cat $O/train.json
python - <<PY
import csv,collections
rows=[r for r in csv.reader(l for l in open("$O/launches_train.csv") if l.startswith('"'))]
h=rows[0]; k=h.index("Kernel Name"); v=h.index("Metric Value")
d=collections.OrderedDict()
for r in rows[1:]:
    n=r[k].split("(")[0][:90]; d.setdefault(n,[0,0.0]); d[n][0]+=1; d[n][1]+=float(r[v].replace(",",""))/1e6
tot=sum(x[1] for x in d.values())
for n,(c,ms) in sorted(d.items(), key=lambda kv:-kv[1][1]): print(f"{ms:9.3f} ms {c:4d}  {n}")
print("TOTAL", tot, "ms", sum(x[0] for x in d.values()), "launches")
PY

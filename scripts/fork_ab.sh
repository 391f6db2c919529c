#!/bin/bash
# development aid: A/B of the forked MLP backwards / env patch (train._step_body) at the 1-GPU batch and at the per-rank shapes of an 8-GPU step
for m in none mlps mlps+env; do
  for w in 0 8; do
    case $m in
      none) export RNERF_FORK_BACKWARD=0;;
      mlps) export RNERF_FORK_BACKWARD=1 RNERF_FORK_ENV=0;;
      *) export RNERF_FORK_BACKWARD=1 RNERF_FORK_ENV=1;;
    esac
    for rep in 1 2; do
      timeout 200 python scripts/train_bench.py --steps 40 --warmup 5 --emulate-world $w 2>/tmp/fork_err.txt | tail -1 > /tmp/fork_ab.json
      python - "$m" "$w" <<'PY'
import json, sys
try:
    d = json.loads(open("/tmp/fork_ab.json").read())
    print(f"fork={sys.argv[1]:10s} world={sys.argv[2]} {d['ms_per_step']:.3f} ms  loss {d['loss']:.6f}", flush=True)
except Exception:
    print("FAILED", sys.argv[1:], open("/tmp/fork_err.txt").read()[-1500:], flush=True)
PY
    done
  done
done

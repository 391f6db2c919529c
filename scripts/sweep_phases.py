"""Development aid: per-phase clock64 stamps of one so3_mlp evaluation per CTA of the reverse sweep (RNERF_SWEEP_PROF).
  RNERF_SWEEP_PROF=/tmp/sweep.txt [RNERF_SWEEP_PROF_EVAL=40] python scripts/train_bench.py --stage all --steps 1 --warmup 0 --eager
  python scripts/sweep_phases.py /tmp/sweep.txt"""
import sys
names = ["entry", "counted", "X+H ready", "raw", "rotation", "dZ4 published", "wgrad D3", "D3 input grads", "D2", "D1", "D0", "enc bwd"]
rows = []
for l in open(sys.argv[1]):
    v = [int(x) for x in l.split()]
    b, n_act, st = v[0], v[1], v[2:]
    t = st[:12]
    if min(t) == 0:
        continue
    rows.append((b, n_act, [t[i + 1] - t[i] for i in range(11)], t[11] - t[0], st[30] - st[29] if st[30] and st[29] else 0))   # st[29], st[30]: start of this / of the next evaluation
print(f"{len(rows)} CTAs reached the stamped evaluation")
rows.sort(key=lambda r: r[1])
for b, n_act, d, tot, loop in rows[:: max(1, len(rows) // 12)]:
    print(f"CTA {b:4d} cols {n_act:3d} total {tot:7d} cyc  eval-to-eval {loop:7d} | " + " ".join(f"{x:6d}" for x in d))
import statistics
print("median per phase: " + " | ".join(f"{names[i + 1]} {int(statistics.median(r[2][i] for r in rows))}" for i in range(11)))
print("median total", int(statistics.median(r[3] for r in rows)), "cycles; median evaluation-to-evaluation", int(statistics.median(r[4] for r in rows if r[4])))

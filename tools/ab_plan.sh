#!/bin/bash
# A/B of the A* kernels on a GPU box (run under gpurun): parity tests on the default build and on the forced-overflow build,
# then bench lines of BASELINE configs 3 and 4 with the small-footprint kernel on / off.  Outputs under gpurun_out/.
T=${1:-ab}
O=gpurun_out
mkdir -p $O
PK=gym_drone2d_activeperception_b200
python -m pytest tests/test_gpu_planner.py tests/test_gpu_fullbatch.py -m gpu -x -q > $O/${T}_tests_small.log 2>&1
echo "tests default build rc=$?"; tail -3 $O/${T}_tests_small.log
if [ -f $PK/libdrone2d_psmall40.so ]; then
D2D_LIB=$PWD/$PK/libdrone2d_psmall40.so python -m pytest tests/test_gpu_planner.py tests/test_gpu_fullbatch.py -m gpu -x -q > $O/${T}_tests_overflow.log 2>&1
echo "tests forced-overflow build rc=$?"; tail -3 $O/${T}_tests_overflow.log
fi
for c in 3 4; do
  python bench.py --config $c --steps 200 --warmup 20 --burn-in 300 --no-cpu-baseline --no-workloads > $O/${T}_bench_cfg${c}_small.json 2> $O/${T}_bench_cfg${c}_small.err
  [ -n "$AB_LARGE" ] && D2D_PLAN_SMALL=0 python bench.py --config $c --steps 200 --warmup 20 --burn-in 300 --no-cpu-baseline --no-workloads > $O/${T}_bench_cfg${c}_large.json 2> $O/${T}_bench_cfg${c}_large.err
  [ -z "$AB_FAST" ] && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${T}_launches_cfg${c}.csv python bench.py --config $c --steps 5 --warmup 3 --burn-in 100 --no-cpu-baseline --no-workloads > $O/${T}_l${c}.log 2>&1
done
T=$T python - <<'PY'
import json, glob, csv, collections, os
T = os.environ["T"]
for f in sorted(glob.glob("gpurun_out/%s_bench_cfg*.json" % T)):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "%.2f M" % (d["value"] / 1e6), "ms %.4f" % d["ms_per_step"], "e2e %.2f M" % (d["e2e"]["value"] / 1e6),
              {k: d["episode_stats"].get(k) for k in ("plans", "plan_failures", "plan_overflows")})
    except Exception as ex:
        print(f, "unreadable", ex)
for f in sorted(glob.glob("gpurun_out/%s_launches_cfg*.csv" % T)):
    agg = collections.defaultdict(list)
    rows = list(csv.reader(l for l in open(f) if l.startswith('"')))
    h = rows[0]; ki = h.index("Kernel Name"); vi = h.index("Metric Value"); ui = h.index("Metric Unit")
    for r in rows[1:]:
        try: agg[r[ki][:48]].append(float(r[vi].replace(",", "")) * (1e-3 if r[ui] == "ns" else 1.0))
        except Exception: pass
    print(f.split("/")[-1])
    for k, v in agg.items():
        v = sorted(v); print("   %-50s n=%3d median %.1f us" % (k, len(v), v[len(v) // 2]))
PY

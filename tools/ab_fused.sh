#!/bin/bash
# A/B of d2d_step_plan_oxford on a GPU box (run under gpurun): planner parity tests, then BASELINE config 4 through the two calls
# (d2d_plan_oxford + d2d_step) and through the fused call with 2 / 3 / 5 A* searches per SM beside the Oxford blocks.
O=gpurun_out; T=${1:-abf}
python -m pytest tests/test_gpu_planner.py -m gpu -x -q > $O/${T}_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/${T}_tests.log
B="python bench.py --config 4 --steps 200 --warmup 20 --burn-in 300 --no-cpu-baseline --no-workloads"
$B --no-fused-oxford > $O/${T}_cfg4_twocalls.json 2> $O/${T}_cfg4_twocalls.err
for s in 2 3 5; do D2D_PLAN_OVERLAP_SLOTS=$s $B > $O/${T}_cfg4_fused_s$s.json 2> $O/${T}_cfg4_fused_s$s.err; done
T=$T python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/%s_cfg4_*.json" % __import__("os").environ["T"])):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); print(f.split("/")[-1], "%.2f M  %.4f ms  e2e %.2f M launches %s" % (d["value"]/1e6, d["ms_per_step"], d["e2e"]["value"]/1e6, d["gpu_launches"]))
    except Exception as ex: print(f, "unreadable", ex)
PY

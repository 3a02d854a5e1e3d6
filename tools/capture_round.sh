#!/bin/bash
# Evidence capture for profiles/ (GPU box, one GPU):  bash tools/capture_round.sh <tag>
# bench lines (all configs + reference arm), ncu launch lists, one `ncu --set full` of the fused step kernel at the
# stationary episode mix, per-warp timeline, sanitizer runs.  Numbers printed under ncu are never bench values.
T=${1:-cap}
O=gpurun_out
mkdir -p $O
(python -m pytest tests -m gpu -q) > $O/${T}_tests.log 2>&1; tail -2 $O/${T}_tests.log
[ -f gym_drone2d_activeperception_b200/libdrone2d_psmall40.so ] && (D2D_LIB=$PWD/gym_drone2d_activeperception_b200/libdrone2d_psmall40.so python -m pytest tests/test_gpu_planner.py tests/test_gpu_fullbatch.py -m gpu -q) > $O/${T}_tests_forced_overflow.log 2>&1; tail -1 $O/${T}_tests_forced_overflow.log
python bench.py --steps 200 --warmup 20 > $O/${T}_bench_default.json 2> $O/${T}_bench_default.err      # headline + workloads list
python bench.py --impl reference --steps 5 --warmup 3 > $O/${T}_bench_cfg2_reference_arm.json 2>> $O/${T}_bench_default.err
python bench.py --config 4 --planner NoMove --gaze scripted --steps 50 --warmup 5 --burn-in 300 --no-cpu-baseline > $O/${T}_bench_cfg4_nomove.json 2> $O/${T}_bench_cfg4.err
python bench.py --config 2 --motion-profile RVO --steps 20 --warmup 5 --burn-in 100 --no-cpu-baseline > $O/${T}_bench_cfg2_rvo.json 2>> $O/${T}_bench_cfg4.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/${T}_launches_cfg2.csv python bench.py --steps 20 --warmup 3 --burn-in 5 --no-cpu-baseline --no-workloads > $O/${T}_l2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${T}_launches_cfg4.csv python bench.py --config 4 --steps 5 --warmup 3 --burn-in 100 --no-cpu-baseline > $O/${T}_l4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:d2d_step_fused_warp_kernel -s 2700 -c 1 -f -o $O/${T}_fused_cfg2 python bench.py --steps 20 --warmup 5 --burn-in 200 --no-cpu-baseline --no-workloads > $O/${T}_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:d2d_oxford_kernel -s 150 -c 1 -f -o $O/${T}_oxford_cfg4 python bench.py --config 4 --steps 5 --warmup 3 --burn-in 150 --no-cpu-baseline > $O/${T}_ncu4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:d2d_step_prim_warp_kernel -s 150 -c 1 -f -o $O/${T}_prim_cfg4 python bench.py --config 4 --steps 5 --warmup 3 --burn-in 150 --no-cpu-baseline > $O/${T}_ncu5.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${T}_launches_cfg3.csv python bench.py --config 3 --steps 5 --warmup 3 --burn-in 100 --no-cpu-baseline --no-workloads > $O/${T}_l3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:d2d_plan_small_kernel -s 150 -c 1 -f -o $O/${T}_plan_small_cfg3 python bench.py --config 3 --steps 5 --warmup 3 --burn-in 150 --no-cpu-baseline --no-workloads > $O/${T}_ncu7.log 2>&1
python tools/plan_prof.py --config 3 --steps 4 > $O/${T}_plan_prof_cfg3.txt 2>&1
python tools/plan_prof.py --config 4 --steps 4 > $O/${T}_plan_prof_cfg4.txt 2>&1
python bench.py --config 4 --steps 200 --warmup 20 --burn-in 300 --no-cpu-baseline --no-workloads --no-fused-oxford > $O/${T}_bench_cfg4_two_calls.json 2>> $O/${T}_bench_cfg4.err
ncu --set full --clock-control none --import-source on -k regex:d2d_rollout_warp_kernel -s 20 -c 1 -f -o $O/${T}_rollout_cfg2 python bench.py --steps 200 --warmup 5 --burn-in 200 --no-cpu-baseline --no-workloads > $O/${T}_ncu6.log 2>&1
python tools/warp_prof.py --config 2 > $O/${T}_warp_timeline_cfg2.txt 2>&1
D2D_PIPE_DEBUG=1 python tools/resident_timeline.py > $O/${T}_resident_timeline_cfg2.txt 2>&1
(for v in 0 1; do echo "D2D_ROLLOUT_VARIANT=$v (0: 28-warp blocks + one block barrier per step, 1: 4-warp blocks, no barrier)"; D2D_ROLLOUT_VARIANT=$v python tools/rollout_bench.py --config 2 --chunks 16,64,200; done) > $O/${T}_rollout_variants_cfg2.txt 2>&1
python tools/rollout_bench.py --config 5 --burn-in 300 --chunks 64 > $O/${T}_rollout_cfg5.txt 2>&1
timeout 300 compute-sanitizer --tool memcheck python tools/sanitize_run.py > $O/${T}_sanitizer_memcheck.txt 2>&1
timeout 300 compute-sanitizer --tool racecheck python tools/sanitize_run.py > $O/${T}_sanitizer_racecheck.txt 2>&1
for f in $O/${T}_bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[1].split('/')[-1], round(d["value"]/1e6,2), "M  ms", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]/1e6,2), "frac", d.get("roofline",{}).get("frac"), "cpu", d.get("cpu_baseline",{}).get("value"))
except Exception as ex: print(sys.argv[1], "unreadable", ex)
PY
done

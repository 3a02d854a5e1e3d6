#!/bin/bash
# Turns the files a `tools/capture_round.sh <tag>` run left under gpurun_out/ into the tracked evidence under profiles/
# (build container, no GPU needed: only reads .ncu-rep / .json / .csv / .txt):  bash tools/publish_profiles.sh <tag> <prefix>
T=${1:-cap}
P=${2:-r2}
O=gpurun_out
D=profiles
for f in bench_default bench_cfg2_reference_arm bench_cfg4_nomove bench_cfg2_rvo bench_cfg4_two_calls; do
  [ -s $O/${T}_$f.json ] && cp $O/${T}_$f.json $D/${P}_$f.json
done
for f in launches_cfg2.csv launches_cfg3.csv launches_cfg4.csv plan_prof_cfg3.txt plan_prof_cfg4.txt warp_timeline_cfg2.txt resident_timeline_cfg2.txt rollout_variants_cfg2.txt rollout_cfg5.txt \
         sanitizer_memcheck.txt sanitizer_racecheck.txt; do
  [ -s $O/${T}_$f ] && cp $O/${T}_$f $D/${P}_$f
done
rep() { [ -s $O/${T}_$1.ncu-rep ] && python tools/ncu_report.py $O/${T}_$1.ncu-rep "$3" > $D/${P}_ncu_$2.txt; }
rep fused_cfg2 fused_cfg2 "ncu --set full --clock-control none --import-source on, one launch of the per-step fused kernel (d2d_step) at BASELINE config 2 (4096 envs, N=10, NoMove), 200 burn-in steps"
rep rollout_cfg2 rollout_cfg2 "ncu --set full, one 200-step launch of d2d_rollout_warp_kernel<28,1,SYNC> at BASELINE config 2 (4096 envs, N=10, NoMove) -- the kernel behind bench.py's headline value"
rep oxford_cfg4 oxford_cfg4 "ncu --set full, d2d_oxford_kernel at BASELINE config 4 (65536 envs, N=24, Primitive + Oxford), 150 burn-in steps"
rep plan_small_cfg3 plan_small_cfg3 "ncu --set full, d2d_plan_small_kernel at BASELINE config 3 (65536 envs, N=142, Primitive planner), 150 burn-in steps"
rep prim_cfg4 prim_cfg4 "ncu --set full, d2d_step_prim_warp_kernel at BASELINE config 4 (65536 envs, N=24, Primitive + Oxford), 150 burn-in steps"
if [ -s $O/${T}_rollout_cfg2.ncu-rep ]; then
  python tools/hot_code.py $O/${T}_rollout_cfg2.ncu-rep d2d_rollout_warp_kernel $((4096*200)) > $D/${P}_hot_code_rollout_cfg2.txt
  python - $O/${T}_rollout_cfg2.ncu-rep $D/traffic_rollout_config2.json <<'PY'
import csv, json, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(raw.splitlines())); hdr, units, d = rows[0], rows[1], rows[2]
def val(name):
    i = hdr.index(name); v = float(d[i].replace(",", "")); u = units[i]
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
b = val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
json.dump({"dram_bytes_per_launch": b, "steps_per_launch": 200,
           "source": "profiles/%s: dram__bytes_read.sum + dram__bytes_write.sum of one 200-step launch of d2d_rollout_warp_kernel at config 2 "
                     "(the env state stays on chip between the steps of a launch: %.2f MB per step against 14.1 MB algorithmic)" %
                     (sys.argv[2].replace("traffic_rollout_config2.json", "").split("/")[-1] + "r2_ncu_rollout_cfg2.txt", b / 200 / 1e6)},
          open(sys.argv[2], "w"))
print(open(sys.argv[2]).read())
PY
fi
[ -s $O/${T}_prim_cfg4.ncu-rep ] && python tools/hot_code.py $O/${T}_prim_cfg4.ncu-rep d2d_step_prim_warp_kernel 65536 > $D/${P}_hot_code_prim_cfg4.txt
ls -la $D | grep " ${P}_" | wc -l

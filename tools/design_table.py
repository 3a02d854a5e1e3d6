"""Rows of DESIGN.md 5.0's table from the bench lines under profiles/ (python tools/design_table.py [prefix])."""
import json, os, sys
P = sys.argv[1] if len(sys.argv) > 1 else "r2"
D = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles")
def row(name, d, cpu=None):
    ssl = d.get("single_step_launches")
    v = "**%.0f M**" % (d["value"] / 1e6) if d["value"] > 5e7 else "%.1f M" % (d["value"] / 1e6)
    if ssl and abs(ssl["value"] / d["value"] - 1.0) > 0.02:
        v += " (K per-step launches: %.0f M)" % (ssl["value"] / 1e6)
    e = d["e2e"]
    ev = "%.0f M" % (e["value"] / 1e6) if e["value"] > 5e7 else "%.1f M" % (e["value"] / 1e6)
    if "full_copy" in e and name.startswith("config 2:"):
        ev = "**%s** (`d2d_step_bound` %.0f M, `d2d_step_host` + mirror %.0f M, plain copies %.0f M)" % (
            ev, e["bound_sync"]["value"] / 1e6, e["mirror_step_host"]["value"] / 1e6, e["full_copy"]["value"] / 1e6)
    fr = d.get("roofline", {}).get("frac")
    print("| %s | %.4f | %s | %s | %s | %s |" % (name, d["ms_per_step"], v, ev, cpu or "—", "%.3f" % fr if fr and d["config"]["planner"] == "NoMove" else "—"))
d = json.load(open(os.path.join(D, P + "_bench_default.json")))
cpu = d.get("cpu_baseline", {}).get("value")
row("config 2: empty map, 4096 envs, N=10, NoMove (headline, `bench.py` default)", d, "%.2f M" % (cpu / 1e6) if cpu else None)
names = ["config 3: random_map_0, 65536 envs, N=142, Primitive planner on device", "config 4: obstacle map, 65536 envs, N=24, Primitive + Oxford on device",
         "config 5: shaped map, 131072 envs, N=96, NoMove", "config 5 corner: 250 rays, FOV 360"]
for n, w in zip(names, d.get("workloads", [])):
    row(n, w)
for f, n in [("_bench_cfg4_nomove.json", "config 4 batch, perception + dynamics only (NoMove)"), ("_bench_cfg2_rvo.json", "config 2 under the RVO motion profile")]:
    p = os.path.join(D, P + f)
    if os.path.isfile(p):
        row(n, json.load(open(p)))

"""CSV rows of the reference's own episode driver (needs /root/reference; build container only).

    python oracle/gen_golden_rows.py

TEST INFRASTRUCTURE ONLY.  Runs the UNMODIFIED `experiment.Experiment(params, dir).run()` (experiment.py:26-106) with
`params.record = True` into a temporary CSV for a few (gaze method, seed) combinations and stores the rows it wrote as
tests/golden/experiment_rows.json: the fixture that pins `experiment.run_batched_rows` / `Experiment.run` of the product."""
import json
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_runner as rr  # noqa: E402
from gen_golden import provenance, OUT  # noqa: E402

CASES = [
    dict(gaze_method="Oxford", planner="Primitive", map_id=1, agent_number=10, agent_radius=15, agent_max_speed=20),
    dict(gaze_method="Oxford", planner="Primitive", map_id=2, agent_number=10, agent_radius=15, agent_max_speed=20),
    dict(gaze_method="LookAhead", planner="Primitive", map_id=3, agent_number=10, agent_radius=15, agent_max_speed=20),
    dict(gaze_method="LookGoal", planner="Primitive", map_id=4, agent_number=10, agent_radius=15, agent_max_speed=20),
    dict(gaze_method="Rotating", planner="Primitive", map_id=5, agent_number=10, agent_radius=15, agent_max_speed=20),
    dict(gaze_method="NoControl", planner="Primitive", map_id=6, agent_number=10, agent_radius=15, agent_max_speed=20),
    dict(gaze_method="Owl", planner="Primitive", map_id=3, agent_number=8),
    dict(gaze_method="Owl", planner="Primitive", map_id=9, agent_number=10, drone_max_speed=20),
    dict(gaze_method="Oxford", planner="Primitive", map_id=0, agent_number=10, agent_radius=10, agent_max_speed=20,
         static_map="maps/obstacle_map.npy"),
]


def main():
    import pandas as pd
    ref_utils, ref_env, _, _ = rr.import_reference()
    import experiment as ref_experiment          # the reference's experiment.py (REFERENCE_ROOT is on sys.path)
    rows = []
    for kw in CASES:
        pk = dict(rr.DEFAULT_PARAMS)
        pk.update(kw)
        params = ref_utils.Params(**pk)
        params.record = True
        with tempfile.TemporaryDirectory() as d, rr._reference_cwd():
            path = os.path.join(d, "rows.csv")
            ref_experiment.Experiment(params, path).run()
            df = pd.read_csv(path, index_col=False)
        assert len(df) == 1, len(df)
        row = {k: (v.item() if hasattr(v, "item") else v) for k, v in df.iloc[0].to_dict().items()}
        rows.append({"params": kw, "row": row})
        print(kw["gaze_method"], kw["map_id"], row["Flight time"], row["Grid discovered"], row["Agent tracked"],
              row["Agent tracked time"], row["Success"], row["Dynamic Collision"])
    with open(os.path.join(OUT, "experiment_rows.json"), "w") as f:
        json.dump({"provenance": provenance(), "rows": rows}, f, indent=1)


if __name__ == "__main__":
    main()

#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
export MOHID_ADT_NO_REBUILD=1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python bench.py --workload c3 --steps 10 --no-e2e --no-cpu-baseline > gpurun_out/b_c3_adapt.json 2> gpurun_out/b_c3_adapt.err; tail -2 gpurun_out/b_c3_adapt.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/b_c3_adapt.json") if l.startswith('{')][0]); print("c3 ms/step %.2f"%d["ms_per_step"], "step_frac %.3f"%d["roofline"]["step_frac"], "launches", d["gpu_launches"], d["checksum"]["total"])
PY

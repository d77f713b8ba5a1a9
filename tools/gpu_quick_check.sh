#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
export MOHID_ADT_NO_REBUILD=1
mkdir -p gpurun_out
timeout 300 python tools/fused_check.py 2>&1 | tail -3
B="timeout 300 python bench.py --workload c3 --steps 5 --no-e2e --no-cpu-baseline"
$B > gpurun_out/b_fused.json 2> gpurun_out/b_fused.err; tail -2 gpurun_out/b_fused.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/b_fused.json")):
    try:
        d=json.loads([l for l in open(f) if l.startswith('{')][0]); print(f, "ms/step %.2f"%d["ms_per_step"], "frac %.3f"%d["roofline"]["frac"], "step_frac %.3f"%d["roofline"]["step_frac"], "launches", d["gpu_launches"], d["checksum"]["total"])
    except Exception as e: print(f, "ERR", e)
PY

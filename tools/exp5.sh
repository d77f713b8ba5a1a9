#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
export MOHID_ADT_NO_REBUILD=1
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
MOHID_ADT_LEAN_ALWAYS=1 python tools/lean_check.py 2>&1 | tail -4
B="python bench.py --workload c3 --steps 5 --no-e2e --no-cpu-baseline"
$B > gpurun_out/b_c3.json 2> gpurun_out/b_c3.err; tail -2 gpurun_out/b_c3.err
MOHID_ADT_PACK_CHUNKED=1 $B > gpurun_out/b_c3_chunked.json 2> gpurun_out/b_c3_chunked.err
MOHID_ADT_NOLEAN=1 $B > gpurun_out/b_c3_old.json 2> gpurun_out/b_c3_old.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/b_c3*.json")):
    try:
        d=json.load(open(f)); print(f, "ms/step %.2f"%d["ms_per_step"], "K2 ms %.2f"%d["roofline"]["kernel_ms"], "frac %.3f"%d["roofline"]["frac"], "launches", d["gpu_launches"])
    except Exception as e: print(f, "ERR", e)
PY

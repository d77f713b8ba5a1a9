#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
export MOHID_ADT_NO_REBUILD=1
mkdir -p gpurun_out; rm -f gpurun_out/b_*.json gpurun_out/b_*.err
python tools/lean_check.py > gpurun_out/lean_check.log 2>&1; tail -2 gpurun_out/lean_check.log
B="python bench.py --workload c3 --steps 5 --no-e2e --no-cpu-baseline"
for d in 0 3; do
  MOHID_ADT_LEAN_WARPS=12 MOHID_ADT_LEAN_PFD=$d $B > gpurun_out/b_lean12_pfd$d.json 2> gpurun_out/b_lean12_pfd$d.err
done
for m in 6 8; do
  MOHID_ADT_COEF_MINB=$m MOHID_ADT_LEAN_WARPS=12 MOHID_ADT_LEAN_PFD=3 $B > gpurun_out/b_lean12_pfd3_minb$m.json 2> gpurun_out/b_minb$m.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/b_*.json")):
    try:
        d=json.load(open(f)); print(f, "ms/step %.2f"%d["ms_per_step"], "K2 ms %.2f"%d["roofline"]["kernel_ms"], "frac %.3f"%d["roofline"]["frac"])
    except Exception as e: print(f, "ERR", e)
PY

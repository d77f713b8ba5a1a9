#!/bin/bash
# round-2 experiment 2: W pack in the look-ahead; selective L2 prefetch in both kernels
cd "$GRAFT_REPO_ROOT" || exit 1
export MOHID_ADT_NO_REBUILD=1
mkdir -p gpurun_out; rm -f gpurun_out/b_*.json
B="python bench.py --workload c3 --steps 5 --no-e2e --no-cpu-baseline"
MOHID_ADT_NOLEAN=1 $B > gpurun_out/b_old.json 2> gpurun_out/b_old.err
MOHID_ADT_NOLEAN=1 MOHID_ADT_PFD=2 $B > gpurun_out/b_old_pfd2.json 2> gpurun_out/b_old_pfd2.err
MOHID_ADT_NOLEAN=1 MOHID_ADT_PFD=4 $B > gpurun_out/b_old_pfd4.json 2> gpurun_out/b_old_pfd4.err
for w in 12 16; do for d in 0 2 3 4; do
  MOHID_ADT_LEAN_WARPS=$w MOHID_ADT_LEAN_PFD=$d $B > gpurun_out/b_lean${w}_pfd$d.json 2> gpurun_out/b_lean${w}_pfd$d.err
done; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/b_*.json")):
    try:
        d=json.load(open(f)); print(f, "ms/step %.2f"%d["ms_per_step"], "K2 ms %.2f"%d["roofline"]["kernel_ms"], "frac %.3f"%d["roofline"]["frac"])
    except Exception as e: print(f, "ERR", e)
PY
python tools/lean_check.py 2>&1 | tail -3

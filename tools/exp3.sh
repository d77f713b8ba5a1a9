#!/bin/bash
# round-2 experiment 3: lean v2 (interleaved packs, direction-free faces, pointer streams)
cd "$GRAFT_REPO_ROOT" || exit 1
export MOHID_ADT_NO_REBUILD=1
mkdir -p gpurun_out; rm -f gpurun_out/b_*.json gpurun_out/b_*.err
python tools/lean_check.py > gpurun_out/lean_check.log 2>&1; tail -4 gpurun_out/lean_check.log
B="python bench.py --workload c3 --steps 5 --no-e2e --no-cpu-baseline"
for w in 12 16; do for d in 0 3; do
  MOHID_ADT_LEAN_WARPS=$w MOHID_ADT_LEAN_PFD=$d $B > gpurun_out/b_lean${w}_pfd$d.json 2> gpurun_out/b_lean${w}_pfd$d.err
done; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/b_*.json")):
    try:
        d=json.load(open(f)); print(f, "ms/step %.2f"%d["ms_per_step"], "K2 ms %.2f"%d["roofline"]["kernel_ms"], "frac %.3f"%d["roofline"]["frac"])
    except Exception as e: print(f, "ERR", e)
PY
MOHID_ADT_LEAN_WARPS=12 MOHID_ADT_LEAN_PFD=3 timeout 600 ncu --set full --import-source on --clock-control none -k regex:lean_kernel -s 2 -c 1 -o gpurun_out/prof_lean12v2 -f $B > gpurun_out/ncu_lean12v2.log 2>&1
MOHID_ADT_LEAN_WARPS=12 timeout 600 ncu --set full --import-source on --clock-control none -k regex:lean_coef -s 1 -c 1 -o gpurun_out/prof_leancoef2 -f $B > gpurun_out/ncu_leancoef2.log 2>&1

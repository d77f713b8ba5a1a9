#!/bin/bash
# round-2 experiment 1: lean kernels: correctness + C3 timing per warps-per-block, ncu of the 16-warp form
cd "$GRAFT_REPO_ROOT" || exit 1
export MOHID_ADT_NO_REBUILD=1
mkdir -p gpurun_out
python tools/lean_check.py > gpurun_out/lean_check.log 2>&1
tail -5 gpurun_out/lean_check.log
MOHID_ADT_NOLEAN=1 python bench.py --workload c3 --steps 5 --no-e2e --no-cpu-baseline > gpurun_out/b_old.json 2> gpurun_out/b_old.err
for w in 8 12 16 20; do
  MOHID_ADT_LEAN_WARPS=$w python bench.py --workload c3 --steps 5 --no-e2e --no-cpu-baseline > gpurun_out/b_lean$w.json 2> gpurun_out/b_lean$w.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/b_*.json")):
    try:
        d=json.load(open(f)); print(f, "ms/step %.2f"%d["ms_per_step"], "K2 ms %.2f"%d["roofline"]["kernel_ms"], "frac %.3f"%d["roofline"]["frac"])
    except Exception as e: print(f, "ERR", e)
PY
MOHID_ADT_LEAN_WARPS=16 timeout 600 ncu --set full --import-source on --clock-control none -k regex:lean_kernel -s 2 -c 1 -o gpurun_out/prof_lean16 -f python bench.py --workload c3 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_lean16.log 2>&1
MOHID_ADT_LEAN_WARPS=16 timeout 600 ncu --set full --clock-control none -k regex:lean_coef -s 1 -c 1 -o gpurun_out/prof_leancoef -f python bench.py --workload c3 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_leancoef.log 2>&1
ls -la gpurun_out | tail -5

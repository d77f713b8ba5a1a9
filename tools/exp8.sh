#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
export MOHID_ADT_NO_REBUILD=1
mkdir -p gpurun_out
cat /sys/fs/cgroup/memory.max /sys/fs/cgroup/memory.current 2>&1
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
timeout 300 python bench.py --workload c2 --steps 20 --no-cpu-baseline > gpurun_out/bench_c2_r2a.json 2> gpurun_out/bench_c2_r2a.err; tail -3 gpurun_out/bench_c2_r2a.err
MOHID_ADT_NOFUSED=1 timeout 300 python bench.py --workload c2 --steps 20 --no-cpu-baseline --no-e2e > gpurun_out/bench_c2_nofused_r2a.json 2> gpurun_out/bench_c2_nofused_r2a.err
timeout 900 python bench.py --steps 10 --no-e2e > gpurun_out/bench_c4_r2a.json 2> gpurun_out/bench_c4_r2a.err; tail -3 gpurun_out/bench_c4_r2a.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/bench_*_r2a.json")):
    try:
        d=json.load(open(f)); print(f, "ms/step %.2f"%d["ms_per_step"], "value %.1f"%d["value"], "frac %.3f"%d["roofline"]["frac"], "step_frac %.3f"%d["roofline"]["step_frac"], "launches", d["gpu_launches"], "e2e", d["e2e"] and round(d["e2e"]["ms_per_step"],1), d.get("e2e_skipped"), "cpu", d["cpu_baseline"] and round(d["cpu_baseline"]["value"],3), "chk", d["checksum"]["total"])
    except Exception as e: print(f, "ERR", e)
PY

#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
export MOHID_ADT_NO_REBUILD=1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_baseline_configs.py -x -q -m gpu -s 2>&1 | tail -15 > gpurun_out/pytest_baseline_cfg.log; cat gpurun_out/pytest_baseline_cfg.log
free -g | head -2
timeout 900 python bench.py --steps 10 > gpurun_out/bench_c4_r2b.json 2> gpurun_out/bench_c4_r2b.err; tail -3 gpurun_out/bench_c4_r2b.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/bench_*_r2b.json")):
    try:
        d=json.load(open(f)); print(f, "ms/step %.2f"%d["ms_per_step"], "value %.1f"%d["value"], "step_frac %.3f"%d["roofline"]["step_frac"], "e2e", d["e2e"] and (round(d["e2e"]["ms_per_step"],1), round(d["e2e"]["value"],2)), d.get("e2e_skipped"), "cpu", d["cpu_baseline"] and round(d["cpu_baseline"]["value"],3), "chk", d["checksum"]["total"])
    except Exception as e: print(f, "ERR", e)
PY

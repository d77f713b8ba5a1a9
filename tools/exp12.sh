#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
export MOHID_ADT_NO_REBUILD=1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_boxes.py tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -5
python bench.py --workload c3 --steps 10 > gpurun_out/bench_c3_r2c.json 2> gpurun_out/bench_c3_r2c.err; tail -3 gpurun_out/bench_c3_r2c.err
python bench.py --workload c3 --steps 10 --bc 4 --no-cpu-baseline > gpurun_out/bench_c3_bc4_r2c.json 2> gpurun_out/bench_c3_bc4_r2c.err; tail -3 gpurun_out/bench_c3_bc4_r2c.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/bench_*_r2c.json")):
    try:
        d=json.loads([l for l in open(f) if l.startswith('{')][0]); print(f, "ms/step %.2f"%d["ms_per_step"], "value %.1f"%d["value"], "step_frac %.3f"%d["roofline"]["step_frac"], "e2e", d["e2e"] and (round(d["e2e"]["ms_per_step"],1), round(d["e2e"]["value"],2)), "res", d.get("e2e_resident") and (round(d["e2e_resident"]["ms_per_step"],1), round(d["e2e_resident"]["value"],2)), "cpu", d["cpu_baseline"] and round(d["cpu_baseline"]["value"],3), "launches", d["gpu_launches"])
    except Exception as e: print(f, "ERR", e)
PY

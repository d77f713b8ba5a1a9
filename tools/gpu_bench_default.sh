#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
export MOHID_ADT_NO_REBUILD=1
mkdir -p gpurun_out
python bench.py --steps 10 > gpurun_out/bench_c4_final.json 2> gpurun_out/bench_c4_final.err; tail -2 gpurun_out/bench_c4_final.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_final.json 2> gpurun_out/bench_ref_final.err
python - <<'PY'
import json
for f in ("gpurun_out/bench_c4_final.json","gpurun_out/bench_ref_final.json"):
    try:
        d=json.loads([l for l in open(f) if l.startswith('{')][0])
        print(f, "ms/step %.2f"%d["ms_per_step"], "value %.3f"%d["value"], d.get("roofline") and ("frac %.3f step_frac %.3f"%(d["roofline"]["frac"], d["roofline"]["step_frac"])), "e2e", d["e2e"] and (round(d["e2e"].get("ms_per_step",0),1), round(d["e2e"]["value"],3)), "res", d.get("e2e_resident") and round(d["e2e_resident"]["ms_per_step"],1), "launches", d["gpu_launches"], "clocks", d.get("clocks"))
    except Exception as e: print(f, "ERR", e)
PY

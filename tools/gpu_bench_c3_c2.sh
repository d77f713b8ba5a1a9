cd "$GRAFT_REPO_ROOT"; export MOHID_ADT_NO_REBUILD=1; mkdir -p gpurun_out
python bench.py --workload c3 --steps 10 > gpurun_out/bench_c3_final.json 2>/dev/null
python bench.py --workload c2 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_c2_final.json 2>/dev/null
python - <<'PY'
import json
for f in ("gpurun_out/bench_c3_final.json","gpurun_out/bench_c2_final.json"):
    d=json.loads([l for l in open(f) if l.startswith('{')][0])
    print(f, "ms/step %.3f"%d["ms_per_step"], "value %.2f"%d["value"], "frac %.3f step_frac %.3f"%(d["roofline"]["frac"], d["roofline"]["step_frac"]), "e2e", d["e2e"] and round(d["e2e"]["ms_per_step"],1), "res", d.get("e2e_resident") and round(d["e2e_resident"]["ms_per_step"],1), "launches", d["gpu_launches"])
PY

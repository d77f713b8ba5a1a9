#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
export MOHID_ADT_NO_REBUILD=1
mkdir -p gpurun_out
for smp in 384x384 1024x1024 2048x1024; do
  python bench.py --impl reference --workload c3 --steps 3 --warmup 1 --ref-sample $smp > gpurun_out/ref_c3_$smp.json 2> gpurun_out/ref_c3_$smp.err
  python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/ref_c3_$smp.json") if l.startswith('{')][0]); print("$smp", round(d["value"],4), d["cpu_baseline"]["sample"])
PY
done
free -g | head -2
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 80 --csv --log-file gpurun_out/launches_r2.csv python bench.py --workload c3 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launch_bench_r2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:adt_transport_fused -s 20 -c 1 -o gpurun_out/prof_fused_c3 -f python bench.py --workload c3 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_fused_c3.log 2>&1
tail -2 gpurun_out/ncu_fused_c3.log | cut -c1-200

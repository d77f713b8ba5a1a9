#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
export MOHID_ADT_NO_REBUILD=1
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:adt_transport_fused -s 3 -c 1 -o gpurun_out/prof_fused_v1 -f \
  python bench.py --workload c3q --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_fused_v1.log 2>&1
tail -3 gpurun_out/ncu_fused_v1.log

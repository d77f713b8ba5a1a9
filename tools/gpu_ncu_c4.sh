#!/bin/bash
# ncu evidence for the default workload (C4, 1 GPU): launch list of the library's kernels and one full capture of the fused kernel
cd "$GRAFT_REPO_ROOT" || exit 1
export MOHID_ADT_NO_REBUILD=1
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:adt_ -s 140 -c 80 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launch_bench_r2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:adt_transport_fused -s 20 -c 1 -o gpurun_out/prof_fused_c4 -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_fused_c4.log 2>&1
tail -1 gpurun_out/ncu_fused_c4.log | cut -c1-200

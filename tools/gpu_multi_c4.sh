#!/bin/bash
# usage: exp10.sh NGPU [extra bench flags]
cd "$GRAFT_REPO_ROOT" || exit 1
export MOHID_ADT_NO_REBUILD=1
N=${1:-2}; shift
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/bench_c4_n$N.json 2> gpurun_out/bench_c4_n$N.err; echo "exit $?"; tail -5 gpurun_out/bench_c4_n$N.err | cut -c1-400
dmesg 2>/dev/null | tail -3
free -g | head -2
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_c4_n$N.json")); print("N=$N ms/step %.2f"%d["ms_per_step"], "value %.1f"%d["value"], "step_frac %.3f"%d["roofline"]["step_frac"], "e2e", d["e2e"] and (round(d["e2e"]["ms_per_step"],1), round(d["e2e"]["value"],2)), d.get("e2e_skipped"), "chk", d["checksum"]["total"], "launches", d["gpu_launches"])
except Exception as e: print("ERR", e)
PY

#!/bin/bash
# usage: exp15.sh NGPU workload...
cd "$GRAFT_REPO_ROOT" || exit 1
export MOHID_ADT_NO_REBUILD=1
N=$1; shift
mkdir -p gpurun_out
for W in "$@"; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --workload $W > gpurun_out/bench_${W}_n$N.json 2> gpurun_out/bench_${W}_n$N.err; echo "exit $?"; tail -4 gpurun_out/bench_${W}_n$N.err | cut -c1-300
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/bench_${W}_n$N.json") if l.startswith('{')][0]); print("$W N=$N ms/step %.2f"%d["ms_per_step"], "value %.1f"%d["value"], "step_frac %.3f"%d["roofline"]["step_frac"], "e2e", d["e2e"] and (round(d["e2e"]["ms_per_step"],1), round(d["e2e"]["value"],2)), d.get("e2e_skipped"), "chk", d["checksum"]["total"], "launches", d["gpu_launches"])
except Exception as e: print("ERR", e)
PY
done
nvidia-smi --query-gpu=memory.used --format=csv | head -3

#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
export MOHID_ADT_NO_REBUILD=1
mkdir -p gpurun_out
rm -f gpurun_out/sanitizer_r2.txt
for tool in memcheck racecheck synccheck; do
  echo "== compute-sanitizer --tool $tool python tools/sanitize_case.py" >> gpurun_out/sanitizer_r2.txt
  timeout 600 compute-sanitizer --tool $tool python tools/sanitize_case.py 2>&1 | grep -E "max rel err|ERROR SUMMARY|RACECHECK SUMMARY|Error|hazard" | head -12 >> gpurun_out/sanitizer_r2.txt
done
cat gpurun_out/sanitizer_r2.txt | cut -c1-200
python bench.py --workload c3 --steps 10 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][0]); print('c3 ms/step %.2f'%d['ms_per_step'], d['checksum']['total'])"

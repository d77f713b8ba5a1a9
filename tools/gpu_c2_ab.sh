cd "$GRAFT_REPO_ROOT"; export MOHID_ADT_NO_REBUILD=1
for v in "" "MOHID_ADT_LEAN_ALWAYS=1"; do
  env $v python bench.py --workload c2 --steps 50 --warmup 5 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][0]); print('$v', 'ms/step %.3f'%d['ms_per_step'], 'step_frac %.3f'%d['roofline']['step_frac'], 'launches', d['gpu_launches'])"
done

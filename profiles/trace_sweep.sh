python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for cfgs in "12 4" "4 4" "8 4" "20 4" "12 2" "12 3" "12 8" "12 1000" "33 1000"; do
  set -- $cfgs
  LB_TRACE_REFILL_MIN=$1 LB_TRACE_TRI_QUARTER=$2 python bench.py --no-cpu-baseline --steps 10 --warmup 4 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['stage_ms']
print('refill $1 triq $2 :', 'ms/frame %.3f'%d['ms_per_step'], ' '.join('%s=%.3f'%(k,v) for k,v in s.items()))"
done

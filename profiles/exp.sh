python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for b in ploc lbvh; do
LB_BVH_BUILDER=$b python bench.py --no-cpu-baseline --steps 10 --warmup 4 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['stage_ms']
print('$b :', 'ms/frame %.3f'%d['ms_per_step'], ' '.join('%s=%.3f'%(k,v) for k,v in s.items()), d['scene'])"
done

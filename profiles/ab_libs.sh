#!/bin/bash
# Usage (GPU box): bash profiles/ab_libs.sh TAG lib1.so lib2.so ...   — A/B of library builds: parity tests + bench stage times per build.
TAG=$1; shift
mkdir -p gpurun_out
for lib in "$@"; do
  export LUMEN_B200_LIB=$PWD/lumenrenderer_b200/$lib
  echo "=== $lib"
  python -m pytest tests/test_gpu_frame.py tests/test_gpu_bsdf.py -m gpu -q 2>&1 | tail -4
  python profiles/parity_report.py 2>&1 | tail -8
  python bench.py --no-cpu-baseline --steps 10 --warmup 4 2>/dev/null | tee gpurun_out/${TAG}_${lib%.so}.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['stage_ms']
print('ms/frame %.3f e2e %.3f'%(d['ms_per_step'], d['e2e']['ms_per_step']), ' '.join('%s=%.3f'%(k,v) for k,v in s.items()))"
done

#!/bin/bash
# Round-1 experiment P (GPU box): full GPU parity suite, then A/B of (a) overlap mode — ReSTIR chain and bounce chain on two streams —
# and its stream priority, (b) builds with the gather kernels / the fused shade kernel compiled for 3 resident blocks per SM.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/p_pytest.log
cat gpurun_out/p_pytest.log
run() {  # tag lib overlap priority
  LUMEN_B200_LIB=$PWD/lumenrenderer_b200/$2 LB_OVERLAP=$3 LB_OVERLAP_PRIORITY=$4 python bench.py --no-cpu-baseline --steps 20 --warmup 5 2>gpurun_out/p_$1.err | tee gpurun_out/p_$1.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['stage_ms']
print('$1: ms/frame %.3f e2e %.3f serial %.3f |'%(d['ms_per_step'], d['e2e']['ms_per_step'], d['overlap']['ms_per_frame_serialised']), ' '.join('%s=%.3f'%(k,v) for k,v in s.items()), d['clocks'])"
}
run base_serial liblumen_b200.so 0 1
run base_overlap_hi liblumen_b200.so 1 1
run base_overlap_lo liblumen_b200.so 1 0
run g3 liblumen_b200_g3.so 1 1
run s3 liblumen_b200_s3.so 1 1
run g3s3 liblumen_b200_g3s3.so 1 1
LUMEN_B200_LIB=$PWD/lumenrenderer_b200/liblumen_b200_g3s3.so python -m pytest tests/test_gpu_frame.py -m gpu -x -q 2>&1 | tail -3

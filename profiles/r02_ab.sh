#!/bin/bash
# Usage (on the GPU box, via gpurun): bash profiles/r02_ab.sh TAG "ENV1=.. ENV2=.." "ENV.." ...
# Runs bench.py (no CPU baseline, 20 steps after 8 warm-up frames) once per environment setting and prints one line per run:
# ms/frame, e2e ms/frame and the exclusive per-stage times. Outputs: gpurun_out/TAG_<k>.json
TAG=$1; shift
mkdir -p gpurun_out
k=0
for envs in "$@"; do
  env $envs python bench.py --no-cpu-baseline --no-extras --steps 20 --warmup 8 > gpurun_out/${TAG}_$k.json 2> gpurun_out/${TAG}_$k.err
  python - "$envs" gpurun_out/${TAG}_$k.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1]); s = d["stage_ms"]
    print(f"[{sys.argv[1]}] ms/frame {d['ms_per_step']:.3f} e2e {d['e2e']['ms_per_step']:.3f} serial {d['overlap']['ms_per_frame_serialised']:.3f} | " + " ".join(f"{k}={v:.3f}" for k, v in s.items()))
except Exception as e:
    print(f"[{sys.argv[1]}] FAILED {e}"); print(open(sys.argv[2].replace('.json', '.err')).read()[-1500:])
PY
  k=$((k+1))
done

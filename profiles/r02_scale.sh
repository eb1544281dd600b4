#!/bin/bash
# Usage (on an N-GPU box, via `gpurun --gpus 8`): bash profiles/r02_scale.sh [MAXN]
# bench.py at N = 1, 2, 4, 8 back to back, launched the way the driver launches it; one JSON line per N in gpurun_out/r02_scale_n<N>.json and
# the device / end-to-end scaling efficiencies (value_N / (N * value_1)) in gpurun_out/r02_scale.json.
MAXN=${1:-8}
mkdir -p gpurun_out
for n in 1 2 4 8; do
  [ $n -gt $MAXN ] && break
  if [ $n -eq 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 20 --warmup 8 --no-cpu-baseline --no-extras > gpurun_out/r02_scale_n$n.json 2> gpurun_out/r02_scale_n$n.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29520 + n)) bench.py --gpus $n --steps 20 --warmup 8 --no-cpu-baseline > gpurun_out/r02_scale_n$n.json 2> gpurun_out/r02_scale_n$n.err
  fi
done
python - $MAXN <<'PY'
import json, sys
rows, base = [], None
for n in (1, 2, 4, 8):
    if n > int(sys.argv[1]): break
    try:
        d = json.loads(open(f"gpurun_out/r02_scale_n{n}.json").read().strip().splitlines()[-1])
    except Exception as e:
        print(n, "FAILED", e, open(f"gpurun_out/r02_scale_n{n}.err").read()[-800:]); continue
    base = base or d
    rows.append({"n_gpus": n, "value": d["value"], "ms_per_step": d["ms_per_step"], "e2e": d["e2e"]["value"], "e2e_ms_per_step": d["e2e"]["ms_per_step"],
                 "efficiency": d["value"] / (n * base["value"]), "e2e_efficiency": d["e2e"]["value"] / (n * base["e2e"]["value"]),
                 "reduce_ms": d.get("reduce_ms"), "extras": {k: (d.get("extras") or {}).get(k) for k in ("strong_bands_ms_per_frame", "c5_time_to_64spp_s")}, "clocks": d.get("clocks")})
    print(rows[-1])
json.dump(rows, open("gpurun_out/r02_scale.json", "w"), indent=1)
PY

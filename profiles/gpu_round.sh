#!/bin/bash
# Usage (on the GPU box, via gpurun): bash profiles/gpu_round.sh TAG [full]
# Runs the GPU parity tests, the bench (not under a profiler), then an ncu launch list of the same bench command and
# optionally one `--set full` capture of a steady-state frame. Outputs go to gpurun_out/TAG_*.
TAG=${1:-x}; FULL=${2:-}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${TAG}_pytest.log
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
if [ -n "$FULL" ]; then
  # skip scene build + 4 frames, then capture one frame's worth of launches
  ncu --set full --clock-control none --import-source on -k regex:'k_(raygen|extend|shade|shadow|fill_bags|ris_order|ris|visibility_shade|temporal|spatial|combine|merge)' --launch-skip 88 -c 22 -f -o gpurun_out/${TAG}_full python bench.py --steps 2 --warmup 5 --no-cpu-baseline > /dev/null 2>&1
fi
cat gpurun_out/${TAG}_pytest.log gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err

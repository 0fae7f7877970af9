#!/bin/bash
# Developer tool (GPU box): tests + native bench + ncu launch list of one step.   usage: tools/gpu_quick.sh <tag>
TAG=${1:-q}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "^E   " | tail -6
timeout 600 python bench.py --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/${TAG}_bench_native.json 2> gpurun_out/${TAG}_bench_native.err
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench_native.json"))
print("value", round(d["value"],4), "e2e", round(d["e2e"]["value"],4), "clocks", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
print({k: round(v*1e3,1) for k,v in d["stage_ms_per_step"].items()})
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_launches.log 2>&1
python tools/launch_list.py gpurun_out/${TAG}_launches.csv

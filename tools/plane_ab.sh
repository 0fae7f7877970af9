#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_plane_gpu.py -x -q 2>&1 | tail -5
for mode in vec scalar; do
  if [ $mode = scalar ]; then export SGS_PLANE_SCALAR=1; else unset SGS_PLANE_SCALAR; fi
  timeout 300 python - <<PY
import json, torch, bench
d = bench.plane_path_timing(torch.device("cuda:0"))
print("$mode", json.dumps({k: (round(v, 4) if isinstance(v, float) else v) for k, v in d.items() if k != "note" and k != "what"}))
PY
done

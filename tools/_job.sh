timeout 900 python -m pytest tests/test_binning_gpu.py tests/test_parity_gpu.py -x -q 2>&1 | tail -2
for i in 1 2; do
timeout 300 python bench.py --steps 60 --warmup 10 --no-cpu-baseline --no-sequence --no-extras 2>gpurun_out/ab_fwd.err > gpurun_out/ab_fwd.json
python - <<PY
import json
d = json.load(open("gpurun_out/ab_fwd.json"))
print("value %.4f e2e %.4f fwd_only %.4f" % (d["value"], d["e2e"]["value"], d["forward_only"]["ms_per_frame"]), d["per_rank"][0]["median_ms"])
PY
done

timeout 900 python -m pytest tests/test_binning_gpu.py -x -q 2>&1 | tail -8
for m in 0 1; do
SGS_BIN_MODE=$m timeout 300 python bench.py --steps 40 --warmup 8 --no-cpu-baseline --no-sequence --no-extras 2>gpurun_out/ab_bin_$m.err > gpurun_out/ab_bin_$m.json
python - <<PY
import json
d = json.load(open("gpurun_out/ab_bin_$m.json"))
s = d["stage_ms_per_step"]
print("bin mode $m value %.4f e2e %.4f fwd_only %.4f" % (d["value"], d["e2e"]["value"], d["forward_only"]["ms_per_frame"]), {k: round(v*1e3,1) for k,v in s.items()})
PY
done
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5

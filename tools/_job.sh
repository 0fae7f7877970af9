for i in 1 2 3; do timeout 600 python -m pytest tests/test_parity_gpu.py -q -s -k "full_size_backward_vs_f64" 2>&1 | grep -E "max|passed|failed|Error" | head -12; done
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 600 python bench.py --steps 50 --warmup 10 --no-sequence 2>gpurun_out/r2i_bench.err > gpurun_out/r2i_bench.json; python - <<PY
import json
d = json.load(open("gpurun_out/r2i_bench.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], "fwd", d["forward_only"]["ms_per_frame"])
print({k: round(v*1e3,1) for k,v in d["stage_ms_per_step"].items()})
print(d["densify_path"])
PY

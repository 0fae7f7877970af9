timeout 900 python -m pytest tests/test_binning_gpu.py tests/test_parity_gpu.py tests/test_abi_c_driver.py -q 2>&1 | tail -3
for pdl in 1 0; do
SGS_NO_PDL=$pdl timeout 300 python bench.py --steps 40 --warmup 8 --no-cpu-baseline --no-sequence --no-extras 2>gpurun_out/ab_pdl_$pdl.err > gpurun_out/ab_pdl_$pdl.json
python - <<PY
import json
d = json.load(open("gpurun_out/ab_pdl_$pdl.json"))
s = d["stage_ms_per_step"]
print("NO_PDL=$pdl value %.4f e2e %.4f fwd_only %.4f" % (d["value"], d["e2e"]["value"], d["forward_only"]["ms_per_frame"]), {k: round(v*1e3,1) for k,v in s.items()})
PY
done

for rep in 1 2; do for pdl in 1 0; do
SGS_NO_PDL=$pdl timeout 300 python bench.py --steps 60 --warmup 10 --no-cpu-baseline --no-sequence --no-extras 2>gpurun_out/ab_pdl_$pdl.err > gpurun_out/ab_pdl_$pdl.json
python - <<PY
import json
d = json.load(open("gpurun_out/ab_pdl_$pdl.json"))
print("NO_PDL=$pdl value %.4f e2e %.4f fwd_only %.4f" % (d["value"], d["e2e"]["value"], d["forward_only"]["ms_per_frame"]))
PY
done; done

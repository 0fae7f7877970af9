#!/bin/bash
# Developer tool (GPU box): A/B of the backward render kernel's reduction variants (SGS_BWD_VARIANT, see
# csrc/sgs_render_bwd.cu::launch_render_bwd) + the gradient parity tests on the default variant.
OUT=gpurun_out
mkdir -p $OUT
for v in ${VARIANTS:-0 1 2 3}; do
  SGS_BWD_VARIANT=$v timeout 300 python bench.py --steps 40 --warmup 8 --no-cpu-baseline --no-sequence --no-extras \
      2>$OUT/ab_bwd_$v.err > $OUT/ab_bwd_$v.json
  python - <<PY
import json
d = json.load(open("$OUT/ab_bwd_$v.json"))
s = d["stage_ms_per_step"]
print("variant $v value %.4f e2e %.4f render_bwd %.1f us render_fwd %.1f us" % (d["value"], d["e2e"]["value"], 1e3 * s["render_bwd"], 1e3 * s["render_fwd"]))
PY
done

#!/bin/bash
# Developer tool (GPU box, via gpurun): everything one round-2 iteration needs from a single GPU slot.
#   usage: tools/gpu_round2.sh [tag] [stages]     stages: any of "quick golden tests bench ncu" (default: all)
TAG=${1:-r2}
STAGES=${2:-"quick golden tests bench ncu train"}
OUT=gpurun_out
mkdir -p $OUT $OUT/golden
has() { [[ " $STAGES " == *" $1 "* ]]; }
if has quick; then
  echo "=== quick: new kernels"
  timeout 600 python -m pytest tests/test_binning_gpu.py tests/test_plane_gpu.py -x -q 2>&1 | tail -15 | tee $OUT/${TAG}_quick.txt
  timeout 300 python tools/binning_phases.py 2>&1 | tail -12 | tee $OUT/${TAG}_phases.txt
fi
if has golden; then
  echo "=== golden: 300-frame sequence from the compiled reference"
  timeout 600 python tests/golden/make_golden.py seq300 2>&1 | tail -3
  [ -f $OUT/golden/config3_seq300.npz ] && cp $OUT/golden/config3_seq300.npz tests/golden/
fi
if has tests; then
  echo "=== pytest -m gpu"
  timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -25 | tee $OUT/${TAG}_pytest.txt
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/${TAG}_smoke.txt
fi
if has bench; then
  echo "=== bench native"
  timeout 900 python bench.py --steps 50 --warmup 10 2> $OUT/${TAG}_bench_native.err > $OUT/${TAG}_bench_native.json
  python - <<PY
import json
d = json.load(open("$OUT/${TAG}_bench_native.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], "fwd", d["forward_only"]["ms_per_frame"], "seq", (d["forward_only"].get("sequence") or {}).get("ms_per_frame"))
print(d.get("stage_ms_per_step"))
print({k: round(v["frac"], 3) for k, v in d.get("hbm_stages", {}).get("stages", {}).items()})
PY
  tail -3 $OUT/${TAG}_bench_native.err
  echo "=== bench reference"
  timeout 900 python bench.py --impl reference --steps 50 --warmup 10 2> $OUT/${TAG}_bench_ref.err > $OUT/${TAG}_bench_reference.json
  python -c "import json;d=json.load(open('$OUT/${TAG}_bench_reference.json'));print('ref value',d['value'],'e2e',d['e2e']['value'],'fwd',d['forward_only']['ms_per_frame'],'seq',(d['forward_only'].get('sequence') or {}).get('ms_per_frame'))"
fi
if has ncu; then
  echo "=== ncu launch list"
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
      --log-file $OUT/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-sequence > $OUT/${TAG}_launches.log 2>&1
  echo "=== ncu full: render kernels"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'render_(fwd|bwd)_kernel' -s 6 -c 2 \
      -f -o $OUT/${TAG}_render python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-sequence > $OUT/${TAG}_ncu_render.log 2>&1
  echo "=== ncu full: binning kernels"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"(binning_fused|tile_count|tile_fill_sorted)_kernel" -s 9 -c 3 \
      -f -o $OUT/${TAG}_binning python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-sequence > $OUT/${TAG}_ncu_binning.log 2>&1
  ls -la $OUT | tail -12
fi
if has train; then
  echo "=== training-time deformation: launch list of one forward + backward, full ncu of the tcgen05 kernels"
  timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file $OUT/${TAG}_deform_train_launches.csv python tools/deform_train_launches.py > $OUT/${TAG}_dt.log 2>&1
  python tools/deform_train_launches.py --summarise $OUT/${TAG}_deform_train_launches.csv | head -12
  timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"deform_(wgrad|train|epilogue)" \
      -f -o $OUT/${TAG}_deform_train python tools/deform_train_launches.py > $OUT/${TAG}_ncu_dt.log 2>&1
  echo "=== plane sampler: launch list"
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv -k regex:"plane_" -c 60 \
      --log-file $OUT/${TAG}_plane_launches.csv python tools/plane_ab.py > $OUT/${TAG}_plane.log 2>&1
  python tools/deform_train_launches.py --summarise $OUT/${TAG}_plane_launches.csv | head -12
fi

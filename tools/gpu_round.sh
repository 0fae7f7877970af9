#!/bin/bash
# Developer tool (GPU box, via gpurun): parity tests, both bench arms, the ncu launch list and a
# full ncu capture of the two compositing kernels.  Everything lands in gpurun_out/.
#   usage: tools/gpu_round.sh [tag]      (tag names the output files, default "run")
TAG=${1:-run}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1

echo "=== pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $OUT/${TAG}_pytest.txt

echo "=== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/${TAG}_smoke.txt

echo "=== bench native"
timeout 600 python bench.py --steps 50 --warmup 10 2> $OUT/${TAG}_bench_native.err | tee $OUT/${TAG}_bench_native.json
echo "=== bench reference"
timeout 600 python bench.py --impl reference --steps 50 --warmup 10 2> $OUT/${TAG}_bench_ref.err | tee $OUT/${TAG}_bench_ref.json

if [ -z "$SKIP_NCU" ]; then
echo "=== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $OUT/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_launches.log 2>&1
echo "=== ncu full: render kernels"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'render_(fwd|bwd)_kernel' -s 6 -c 2 \
    -f -o $OUT/${TAG}_render python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_render.log 2>&1
ls -la $OUT
fi
if [ -n "$EXTRA_CMD" ]; then
    echo "=== extra: $EXTRA_CMD"
    bash -c "$EXTRA_CMD" 2>&1 | tail -40
fi

#!/bin/bash
# Developer tool (GPU box with 8 GPUs, via `gpurun --gpus 8`): the N = 8 legs only (the N = 1 / 2 legs are in profiles/r2n_*).
N=${1:-8}
OUT=gpurun_out
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
echo "=== bench N=$N"
timeout 400 $TR --nproc-per-node $N --master-port 29538 bench.py --gpus $N --steps 30 --warmup 8 --no-sequence 2>$OUT/multi_b_$N.err | tail -1 > $OUT/multi_bench_n$N.json
python -c "import json;d=json.load(open('$OUT/multi_bench_n$N.json'));print(d['n_gpus'],d['value'],d['ms_per_step'],d['e2e']['value'],[ (r['rank'],round(r['median_ms'],4)) for r in d['per_rank']])"
echo "=== dp_train N=$N"
timeout 400 $TR --nproc-per-node $N --master-port 29518 tools/dp_train.py --iters 20 2>$OUT/multi_dp_$N.err | tail -1 | tee $OUT/multi_dp_train_n$N.json
echo "=== config5 N=$N"
timeout 500 $TR --nproc-per-node $N --master-port 29512 tools/config5_render.py 2>$OUT/multi_c5_n.err | tail -1 | tee $OUT/multi_config5_n$N.json

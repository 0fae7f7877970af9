#!/bin/bash
# Developer tool (GPU box with N GPUs, via `gpurun --gpus N`): the multi-GPU legs the single-GPU tiers cannot run.
#   configs[4] (6 000 renders sharded over the ranks, metric sums independent of N), data-parallel training
#   iterations with the overlapped gradient exchange, and the weak-scaling bench.  Output: gpurun_out/multi_*.json
N=${1:-8}
OUT=gpurun_out
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
echo "=== config5 N=1"; timeout 600 python tools/config5_render.py 2>$OUT/multi_c5_1.err | tail -1 | tee $OUT/multi_config5_n1.json
echo "=== config5 N=$N"; timeout 600 $TR --nproc-per-node $N --master-port 29512 tools/config5_render.py 2>$OUT/multi_c5_n.err | tail -1 | tee $OUT/multi_config5_n$N.json
for n in 1 2 4 $N; do
  [ $n -gt $N ] && continue
  echo "=== dp_train N=$n"
  if [ $n -eq 1 ]; then timeout 600 python tools/dp_train.py --iters 20 2>$OUT/multi_dp_$n.err | tail -1 | tee $OUT/multi_dp_train_n$n.json
  else timeout 600 $TR --nproc-per-node $n --master-port 2951$n tools/dp_train.py --iters 20 2>$OUT/multi_dp_$n.err | tail -1 | tee $OUT/multi_dp_train_n$n.json
       timeout 600 $TR --nproc-per-node $n --master-port 2952$n tools/dp_train.py --iters 20 --no-overlap 2>>$OUT/multi_dp_$n.err | tail -1 | tee $OUT/multi_dp_train_n${n}_noverlap.json; fi
done
for n in 1 2 4 $N; do
  [ $n -gt $N ] && continue
  echo "=== bench N=$n"
  if [ $n -eq 1 ]; then timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-sequence 2>$OUT/multi_b_$n.err | tail -1 > $OUT/multi_bench_n$n.json
  else timeout 600 $TR --nproc-per-node $n --master-port 2953$n bench.py --gpus $n --steps 20 --warmup 5 --no-sequence 2>$OUT/multi_b_$n.err | tail -1 > $OUT/multi_bench_n$n.json; fi
  python -c "import json;d=json.load(open('$OUT/multi_bench_n$n.json'));print(d['n_gpus'],d['ms_per_step'],d['e2e']['value'],[ (r['rank'],round(r['median_ms'],4)) for r in d['per_rank']])"
done

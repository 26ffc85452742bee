#!/bin/bash
# 1/2/4/8-GPU runs of bench.py and of the training driver on one 8-GPU box.  The 4-, 2- and 1-GPU
# runs use disjoint GPUs and run side by side; the 8-GPU run has the box to itself.
# usage: tools/scaling.sh <workload> <outdir>
W=${1:-ogbn-products}; OUT=${2:-gpurun_out/scaling}; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
bench() { # n gpus port
  if [ $1 -eq 1 ]; then CUDA_VISIBLE_DEVICES=$2 python bench.py --workload $W --steps 20 --warmup 5 --no-cpu-baseline
  else CUDA_VISIBLE_DEVICES=$2 $TR --nproc-per-node $1 --master-port $3 bench.py --gpus $1 --workload $W --steps 20 --warmup 5; fi; }
train() {
  if [ $1 -eq 1 ]; then CUDA_VISIBLE_DEVICES=$2 python -m spgemm_gnn_b200.train --dataset $W --model sage --epochs 10 --norm
  else CUDA_VISIBLE_DEVICES=$2 $TR --nproc-per-node $1 --master-port $3 -m spgemm_gnn_b200.train --dataset $W --model sage --epochs 10 --norm; fi; }
bench 4 0,1,2,3 29601 > $OUT/bench_$W.4.log 2>&1 &
bench 2 4,5 29602 > $OUT/bench_$W.2.log 2>&1 &
bench 1 6 0 > $OUT/bench_$W.1.log 2>&1 &
wait
bench 8 0,1,2,3,4,5,6,7 29603 > $OUT/bench_$W.8.log 2>&1
train 4 0,1,2,3 29604 > $OUT/train_$W.4.log 2>&1 &
train 2 4,5 29605 > $OUT/train_$W.2.log 2>&1 &
train 1 6 0 > $OUT/train_$W.1.log 2>&1 &
wait
train 8 0,1,2,3,4,5,6,7 29606 > $OUT/train_$W.8.log 2>&1
for n in 1 2 4 8; do
  tail -1 $OUT/bench_$W.$n.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bench $W gpus', d['n_gpus'], 'ms/step %.3f' % d['ms_per_step'], 'edges/s %.3e' % d['value'], 'fwd %.3f bwd %.3f' % (d['kernels']['spgemm_fwd_ms'], d['kernels']['sspmm_bwd_ms']), 'e2e %.2f ms' % d['e2e']['ms_per_step'])" 2>&1 | tail -1
  tail -1 $OUT/train_$W.$n.log
done

#!/bin/bash
# Training epoch time at 1/2/4/8 GPUs for one (dataset, model, maxk); 4/2/1 side by side, then 8.
# usage: tools/scaling_train.sh <dataset> <model> <maxk> <outdir> [ns="1 2 4 8"]
W=$1; M=$2; K=$3; OUT=${4:-gpurun_out/scaling}; NS=${5:-"1 2 4 8"}; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
ARGS="--dataset $W --model $M --maxk $K --epochs 8 --norm"
run() { # n gpus port
  if [ $1 -eq 1 ]; then CUDA_VISIBLE_DEVICES=$2 python -m spgemm_gnn_b200.train $ARGS
  else CUDA_VISIBLE_DEVICES=$2 $TR --nproc-per-node $1 --master-port $3 -m spgemm_gnn_b200.train $ARGS; fi; }
P=$((29700 + RANDOM % 200))
for n in $NS; do case $n in
  4) run 4 0,1,2,3 $((P+1)) > $OUT/train_${W}_${M}_k$K.4.log 2>&1 & ;;
  2) run 2 4,5 $((P+2)) > $OUT/train_${W}_${M}_k$K.2.log 2>&1 & ;;
  1) run 1 6 0 > $OUT/train_${W}_${M}_k$K.1.log 2>&1 & ;;
esac; done
wait
for n in $NS; do [ $n -eq 8 ] && run 8 0,1,2,3,4,5,6,7 $((P+3)) > $OUT/train_${W}_${M}_k$K.8.log 2>&1; done
for n in $NS; do tail -1 $OUT/train_${W}_${M}_k$K.$n.log; done

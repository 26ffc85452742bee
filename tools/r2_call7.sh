#!/bin/bash
# round 2, GPU call 7 (8 GPUs): forms of the forward exchange -- copy engines on 1 / 7 side streams,
# SpGEMM waiting per block or for the whole table -- against the NCCL form, Reddit and products shapes.
OUT=gpurun_out/r2; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
i=0
for cfg in "1 1" "7 1" "7 0" "3 1"; do
  set -- $cfg; i=$((i+1))
  echo "== MAXK_PEER_STREAMS=$1 MAXK_PEER_OVERLAP=$2" >> $OUT/exchange_forms8.log
  MAXK_PEER_STREAMS=$1 MAXK_PEER_OVERLAP=$2 timeout 300 $TR --nproc-per-node 8 --master-port $((29680+i)) tools/peer_check.py dist --bench --products 2>&1 \
    | grep -v '^\*\|OMP_NUM\|^W1\|^$\|NCCL version' >> $OUT/exchange_forms8.log
done
cat $OUT/exchange_forms8.log

#!/bin/bash
# round 2, GPU call 21 (2 GPUs): push forms at 2 ranks (multicast, stand-alone NVLink stores, pusher CTAs, copy engines, auto).
OUT=gpurun_out/r2; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
export MAXK_PEER_TIMEOUT_MS=15000
timeout 420 $TR --nproc-per-node 2 --master-port 29721 tools/peer_check.py dist --bench --products --sweep 2>&1 \
  | grep -v '^\*\|OMP_NUM\|^W1\|^$\|NCCL version' > $OUT/peer_forms2.log; echo "sweep rc=${PIPESTATUS[0]}"
cat $OUT/peer_forms2.log

#!/bin/bash
OUT=gpurun_out/r2; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 2 --master-port 29663 tools/push_probe.py > $OUT/push_probe.log 2>&1; echo rc=$?
grep -v '^\*\|OMP_NUM\|^W1' $OUT/push_probe.log | tail -20

#!/bin/bash
# round 2, GPU call 2 (1 GPU): new peer protocol in the single-process emulation, parity suite,
# record order (longest first vs row order) and backward record size at full scale and on an 8-way shard.
OUT=gpurun_out/r2; mkdir -p $OUT
timeout 600 python tools/peer_check.py virtual 4 --all-k > $OUT/peer_virtual.log 2>&1; echo "virtual rc=$?"; tail -12 $OUT/peer_virtual.log
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest2.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest2.log
for eo in 0 1; do for s in 1 8; do
  MAXK_EXEC_ORDER=$eo timeout 300 python tools/variant_bench.py --shard $s --ks 32 --tag exec$eo
done; done > $OUT/exec_order.log 2>&1
for mz in 256 512 1024; do for s in 1 8; do
  MAXK_BWD_MAX_NZ=$mz timeout 300 python tools/variant_bench.py --shard $s --ks 32 --tag bwdnz$mz
done; done >> $OUT/exec_order.log 2>&1
MAXK_EXEC_ORDER=1 timeout 300 python tools/variant_bench.py --workload ogbn-products --ks 32 --tag exec1 >> $OUT/exec_order.log 2>&1
MAXK_EXEC_ORDER=0 timeout 300 python tools/variant_bench.py --workload ogbn-products --ks 32 --tag exec0 >> $OUT/exec_order.log 2>&1
MAXK_EXEC_ORDER=1 timeout 300 python tools/variant_bench.py --workload ogbn-products --shard 8 --ks 32 --tag exec1 >> $OUT/exec_order.log 2>&1
MAXK_EXEC_ORDER=0 timeout 300 python tools/variant_bench.py --workload ogbn-products --shard 8 --ks 32 --tag exec0 >> $OUT/exec_order.log 2>&1
cat $OUT/exec_order.log

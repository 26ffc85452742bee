#!/bin/bash
OUT=gpurun_out/r2; mkdir -p $OUT
timeout 600 python tools/peer_check.py virtual 4 --all-k > $OUT/peer_virtual8.log 2>&1; echo "virtual rc=$?"; grep -c OK $OUT/peer_virtual8.log; tail -4 $OUT/peer_virtual8.log
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest8.log 2>&1; echo "pytest rc=$?"; tail -6 $OUT/pytest8.log
for mb in 0 32 48 64 96; do
  if [ $mb = 0 ]; then export MAXK_BWD_TILED=0; else export MAXK_BWD_TILED=auto MAXK_BWD_TILE_MB=$mb; fi
  timeout 300 python tools/variant_bench.py --workload ogbn-products --ks 32 --tag tile$mb
  timeout 300 python tools/variant_bench.py --workload ogbn-products --shard 8 --ks 32 --tag tile$mb
done > $OUT/bwd_tiled.log 2>&1
unset MAXK_BWD_TILED MAXK_BWD_TILE_MB
MAXK_BWD_TILED=1 MAXK_BWD_TILE_MB=16 timeout 300 python tools/variant_bench.py --workload reddit --ks 32 --tag tile16_forced >> $OUT/bwd_tiled.log 2>&1
cat $OUT/bwd_tiled.log

#!/bin/bash
# round 2, GPU call 27 (1 GPU): tile size of the column-blocked backward on the products shape (and an 8-way shard of it).
OUT=gpurun_out/r2; mkdir -p $OUT
{
for mb in 64 80 96 112 128 160; do
  MAXK_BWD_TILED=auto MAXK_BWD_TILE_MB=$mb timeout 300 python tools/variant_bench.py --workload ogbn-products --ks 32 --tag tile$mb
  MAXK_BWD_TILED=auto MAXK_BWD_TILE_MB=$mb timeout 300 python tools/variant_bench.py --workload ogbn-products --shard 8 --ks 32 --tag tile$mb
done
MAXK_BWD_TILED=auto MAXK_BWD_TILE_MB=96 timeout 300 python tools/variant_bench.py --workload ogbn-products --ks 16,64 --tag tile96
MAXK_BWD_TILED=auto MAXK_BWD_TILE_MB=64 timeout 300 python tools/variant_bench.py --workload ogbn-products --ks 16,64 --tag tile64
} > $OUT/bwd_tile_sizes.log 2>&1
cat $OUT/bwd_tile_sizes.log

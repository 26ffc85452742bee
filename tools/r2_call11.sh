#!/bin/bash
# round 2, GPU call 11 (1 GPU): second-generation top-k kernel + fused banking: parity, then timing against
# the first kernel (MAXK_TOPK_V1=1) on the Reddit / products / Flickr shapes and k = 8..64.
OUT=gpurun_out/r2; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest11.log 2>&1; echo "pytest rc=$?"; tail -15 $OUT/pytest11.log
{
for v in 0 1; do
  MAXK_TOPK_V1=$v timeout 300 python tools/variant_bench.py --workload reddit --ks 8,16,32,64 --topk --tag topk_v1=$v
  MAXK_TOPK_V1=$v timeout 300 python tools/variant_bench.py --workload flickr --ks 32 --topk --tag topk_v1=$v
  MAXK_TOPK_V1=$v timeout 300 python tools/variant_bench.py --workload ogbn-products --ks 32 --topk --tag topk_v1=$v
done
} > $OUT/topk_tile.log 2>&1
cat $OUT/topk_tile.log

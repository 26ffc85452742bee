#!/bin/bash
# round 2, GPU call 17 (1 GPU): full GPU suite again; records per CTA of the backward (MK_BWD_WARPS = 1, 2, 4).
OUT=gpurun_out/r2; mkdir -p $OUT
timeout 1800 python -m pytest tests -m gpu -x -q -s > $OUT/pytest17.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|flickr shape|: loss " $OUT/pytest17.log | tail -12
{
timeout 300 python tools/variant_bench.py --workload reddit --ks 8,32,64 --tag bwd_warps1
for w in 2 4; do
  MAXK_LIB=$PWD/spgemm_gnn_b200/libmaxk_bwdw$w.so timeout 300 python tools/variant_bench.py --workload reddit --ks 8,32,64 --tag bwd_warps$w
done
timeout 300 python tools/variant_bench.py --workload ogbn-products --ks 32 --tag bwd_warps1
MAXK_BWD_TILED=0 timeout 300 python tools/variant_bench.py --workload ogbn-products --ks 32 --tag bwd_warps1_plain
MAXK_BWD_TILED=0 MAXK_LIB=$PWD/spgemm_gnn_b200/libmaxk_bwdw2.so timeout 300 python tools/variant_bench.py --workload ogbn-products --ks 32 --tag bwd_warps2_plain
} > $OUT/bwd_warps_per_cta.log 2>&1
cat $OUT/bwd_warps_per_cta.log

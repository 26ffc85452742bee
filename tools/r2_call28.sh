#!/bin/bash
# round 2, GPU call 28 (1 GPU): group-private banks in the plain forward (short-record shapes) against the
# end-to-end copies it replaces; parity suite.
OUT=gpurun_out/r2; mkdir -p $OUT
timeout 1800 python -m pytest tests -m gpu -x -q > $OUT/pytest28.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest28.log
{
for w in ogbn-products flickr yelp; do
  K=32; D=256; if [ $w = yelp ]; then K=16; D=384; fi
  timeout 300 python tools/variant_bench.py --workload $w --ks $K --dim $D --tag group_banks
  MAXK_LIB=$PWD/spgemm_gnn_b200/libmaxk_vecold.so timeout 300 python tools/variant_bench.py --workload $w --ks $K --dim $D --tag copies_end_to_end
done
timeout 300 python tools/variant_bench.py --workload ogbn-products --ks 8,16,64 --tag group_banks
MAXK_LIB=$PWD/spgemm_gnn_b200/libmaxk_vecold.so timeout 300 python tools/variant_bench.py --workload ogbn-products --ks 8,16,64 --tag copies_end_to_end
timeout 300 python tools/variant_bench.py --workload ogbn-products --shard 8 --ks 32 --tag group_banks
MAXK_LIB=$PWD/spgemm_gnn_b200/libmaxk_vecold.so timeout 300 python tools/variant_bench.py --workload ogbn-products --shard 8 --ks 32 --tag copies_end_to_end
} > $OUT/fwd_vec_group_banks.log 2>&1
cat $OUT/fwd_vec_group_banks.log

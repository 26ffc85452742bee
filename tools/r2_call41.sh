#!/bin/bash
# round 2, call 41: steps in flight of the backward (MK_BWD_U) and of the plain forward (MK_FWD_VEC_U) per width and shape.
OUT=$PWD/gpurun_out/r2; mkdir -p $OUT; : > $OUT/bwd_fwdvec_unroll.log
for v in b200 bu2 bu4 bu16; do
  MAXK_LIB=$PWD/spgemm_gnn_b200/libmaxk_$v.so timeout 300 python tools/variant_bench.py --ks 8,16,32,64 >> $OUT/bwd_fwdvec_unroll.log 2>&1
done
for v in b200 bu4 bu16 fu2 fu8; do
  MAXK_LIB=$PWD/spgemm_gnn_b200/libmaxk_$v.so timeout 300 python tools/variant_bench.py --workload ogbn-products --ks 32 >> $OUT/bwd_fwdvec_unroll.log 2>&1
done
grep -v Warn $OUT/bwd_fwdvec_unroll.log | cut -c1-200

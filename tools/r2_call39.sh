#!/bin/bash
# round 2, call 39: neighbour steps in flight (U) of the banked forward per width, after the per-width instantiations.
OUT=$PWD/gpurun_out/r2; mkdir -p $OUT; : > $OUT/fwd_unroll.log
for v in b200 u2 u4 u8; do
  MAXK_LIB=$PWD/spgemm_gnn_b200/libmaxk_$v.so timeout 300 python tools/variant_bench.py --ks 8,16,32,64 >> $OUT/fwd_unroll.log 2>&1
done
grep -v Warn $OUT/fwd_unroll.log | cut -c1-200

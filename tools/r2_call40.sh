#!/bin/bash
# round 2, call 40: with U(16) = 4: plain forward on the bare / combined instantiation per width, LayerNorm form with /
# without prefetch per width.
OUT=$PWD/gpurun_out/r2; mkdir -p $OUT; : > $OUT/fwd_u4_masks.log
for v in b200 bare0 bareall; do
  MAXK_LIB=$PWD/spgemm_gnn_b200/libmaxk_$v.so timeout 300 python tools/variant_bench.py --ks 8,16,32,64 >> $OUT/fwd_u4_masks.log 2>&1
done
for v in b200 pfall; do
  MAXK_LIB=$PWD/spgemm_gnn_b200/libmaxk_$v.so timeout 300 python tools/ln_epilogue_bench.py reddit 8,16,32,64 >> $OUT/fwd_u4_masks.log 2>&1
done
grep -v Warn $OUT/fwd_u4_masks.log | cut -c1-250

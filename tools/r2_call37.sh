#!/bin/bash
# round 2, call 37: __launch_bounds__(32, MINBLOCKS) on the banked forward: 1 (compiler's choice, shipped) / 20 / 25 / 32.
OUT=$PWD/gpurun_out/r2; mkdir -p $OUT; : > $OUT/fwd_minblocks.log
for v in b200 mb20 mb25 mb32; do
  MAXK_LIB=$PWD/spgemm_gnn_b200/libmaxk_$v.so timeout 300 python tools/variant_bench.py --ks 8,16,32,64 >> $OUT/fwd_minblocks.log 2>&1
done
grep -v Warn $OUT/fwd_minblocks.log | cut -c1-200

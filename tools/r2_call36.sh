#!/bin/bash
# round 2, call 36: forward with the epilogue as its own instantiation (no h_self registers in the plain launches);
# h_self prefetch at CTA start per width: default (8, 32, 64), all widths, none.
OUT=$PWD/gpurun_out/r2; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "forward or layernorm or epilogue or banked or packed or fused or reddit" > $OUT/pytest36.log 2>&1
echo "pytest rc=$?"; tail -2 $OUT/pytest36.log
{
timeout 300 python tools/variant_bench.py --ks 8,16,32,64 --tag default
for v in default pf_all pf_none; do
  L=$PWD/spgemm_gnn_b200/libmaxk_$v.so; [ $v = default ] && L=$PWD/spgemm_gnn_b200/libmaxk_b200.so
  MAXK_LIB=$L timeout 300 python tools/ln_epilogue_bench.py reddit 8,16,32,64
done
} > $OUT/epi_prefetch.log 2>&1
grep -v Warn $OUT/epi_prefetch.log | cut -c1-250

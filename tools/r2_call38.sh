#!/bin/bash
# round 2, call 38: per-width choice of the forward instantiation (plain: k = 16 without the epilogue code, 8 / 32 / 64
# with it; LayerNorm form: prefetch at 8 / 32 only): parity, then the plain and LayerNorm-form timings.
OUT=$PWD/gpurun_out/r2; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_models.py -x -q -m gpu -k "not flickr_shape and not fifty" > $OUT/pytest38.log 2>&1
echo "pytest rc=$?"; tail -2 $OUT/pytest38.log
{ timeout 300 python tools/variant_bench.py --ks 8,16,32,64 --tag final
  timeout 300 python tools/ln_epilogue_bench.py reddit 8,16,32,64; } > $OUT/fwd_per_width.log 2>&1
grep -v Warn $OUT/fwd_per_width.log | cut -c1-250

#!/bin/bash
# round 2, GPU call 15 (1 GPU): parity; the LayerNorm epilogue inside the forward (h_self prefetched)
# against the separate kernels; MaxK-SAGE epoch with / without it and with / without MAXK_ALIGN_GEMM.
OUT=gpurun_out/r2; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest15.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest15.log
timeout 300 python tools/ln_epilogue_bench.py reddit > $OUT/ln_epilogue.log 2>&1; cat $OUT/ln_epilogue.log
for cfg in "1 0" "0 0" "1 1"; do
  set -- $cfg
  MAXK_FUSED_LN=$1 MAXK_ALIGN_GEMM=$2 timeout 600 python tools/epoch_profile.py --model maxk-sage --tf32 > $OUT/epoch_ln$1_align$2.txt 2>&1
  echo "fused_ln=$1 align_gemm=$2 rc=$? $(grep 'epoch ms' $OUT/epoch_ln$1_align$2.txt)"
done

#!/bin/bash
# round 2, call 50: bank assignment with bit-sliced membership masks (same assignment, fewer instructions).
OUT=$PWD/gpurun_out/r2; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_properties.py -x -q -m gpu -k "bank or packed or forward or fused or topk_tile" > $OUT/pytest50.log 2>&1
echo "pytest rc=$?"; tail -2 $OUT/pytest50.log
timeout 300 python tools/variant_bench.py --ks 8,16,32,64 --topk --tag bank2 > $OUT/bank2.log 2>&1
grep -v Warn $OUT/bank2.log | cut -c1-260

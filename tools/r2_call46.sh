#!/bin/bash
# round 2, call 46: fold kernel with one thread per record for the detection and float4 folds.
OUT=$PWD/gpurun_out/r2; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_properties.py tests/test_gpu_binding.py -x -q -m gpu > $OUT/pytest46.log 2>&1
echo "pytest rc=$?"; tail -2 $OUT/pytest46.log
{ timeout 300 python tools/variant_bench.py --ks 8,16,32,64 --tag fold2
  timeout 300 python tools/variant_bench.py --workload ogbn-proteins --ks 64 --tag fold2
  timeout 300 python tools/ln_epilogue_bench.py reddit 32; } > $OUT/fold2.log 2>&1
grep -v Warn $OUT/fold2.log | cut -c1-250

#!/bin/bash
# round 2, call 30: lane-contiguous top-k kernel (one prefix sum, shared-memory compaction, 256-bit loads,
# lane-maximum lower bound) against the strided round-1 mapping; parity first.
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_properties.py -x -q -m gpu -k "topk or scatter or golden or maxk or propert or cbsr" > gpurun_out/r2/pytest30.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2/pytest30.log
timeout 300 python tools/topk_ab.py lane > gpurun_out/r2/topk_ab30.log 2>&1
MAXK_TOPK_STRIDED=1 timeout 300 python tools/topk_ab.py strided >> gpurun_out/r2/topk_ab30.log 2>&1
cat gpurun_out/r2/topk_ab30.log

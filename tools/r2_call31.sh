#!/bin/bash
# round 2, call 31+: top-k after each tweak (parity, timing, one ncu capture of the lane kernel).
mkdir -p gpurun_out/r2
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_properties.py -x -q -m gpu -k "topk or propert" > gpurun_out/r2/pytest31.log 2>&1
echo "pytest rc=$?"; tail -2 gpurun_out/r2/pytest31.log
timeout 300 python tools/topk_ab.py lane > gpurun_out/r2/topk_ab31.log 2>&1; cat gpurun_out/r2/topk_ab31.log
timeout 200 python tools/topk_ab.py lane 256:32 > gpurun_out/r2/topk_plain31.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:topk_cbsr -s 5 -c 1 -f -o gpurun_out/r2/prof_r2_topk_lane python tools/topk_ab.py lane 256:32 > gpurun_out/r2/ncu_topk31.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/r2/ncu_topk31.log

#!/bin/bash
# round 2, GPU call 3 (1 GPU): packed banked forward at k = 8, 16 (parity + timing against the plain kernels)
OUT=gpurun_out/r2; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "packed or arrival or banked" > $OUT/pytest3.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest3.log
for pk in 0 1; do MAXK_PACKED=$pk timeout 300 python tools/variant_bench.py --ks 8,16 --tag packed$pk; done > $OUT/packed.log 2>&1
MAXK_PACKED=1 timeout 300 python tools/variant_bench.py --workload ogbn-proteins --ks 8,16,64 --tag packed1 >> $OUT/packed.log 2>&1
MAXK_PACKED=0 timeout 300 python tools/variant_bench.py --workload ogbn-proteins --ks 8,16 --tag packed0 >> $OUT/packed.log 2>&1
cat $OUT/packed.log

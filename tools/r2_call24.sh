#!/bin/bash
# round 2, GPU call 24 (1 GPU): what a third copy of the accumulator cells (12 KB instead of 8 KB per CTA:
# 17 instead of 25 resident CTAs) would cost the banked forward before it gains anything.
OUT=gpurun_out/r2; mkdir -p $OUT
{
timeout 300 python tools/variant_bench.py --workload reddit --ks 8,32,64 --tag smem8k
MAXK_LIB=$PWD/spgemm_gnn_b200/libmaxk_smem12.so timeout 300 python tools/variant_bench.py --workload reddit --ks 8,32,64 --tag smem12k
} > $OUT/fwd_smem_occupancy.log 2>&1
cat $OUT/fwd_smem_occupancy.log

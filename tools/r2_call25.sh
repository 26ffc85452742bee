#!/bin/bash
# round 2, GPU call 25 (1 GPU): the backward with FEWER resident CTAs (32 / 24 / 15 per SM, capped by dummy shared memory).
OUT=gpurun_out/r2; mkdir -p $OUT
{
timeout 300 python tools/variant_bench.py --workload reddit --ks 8,32,64 --tag bwd_32ctas
MAXK_LIB=$PWD/spgemm_gnn_b200/libmaxk_bwdsm7168.so timeout 300 python tools/variant_bench.py --workload reddit --ks 8,32,64 --tag bwd_24ctas
MAXK_LIB=$PWD/spgemm_gnn_b200/libmaxk_bwdsm12288.so timeout 300 python tools/variant_bench.py --workload reddit --ks 8,32,64 --tag bwd_15ctas
} > $OUT/bwd_residency.log 2>&1
cat $OUT/bwd_residency.log

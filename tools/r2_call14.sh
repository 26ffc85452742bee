#!/bin/bash
# round 2, GPU call 14 (1 GPU): the ncu evidence of the round -- launch list of bench.py, then one
# `--set full` capture per kernel family (each only after the same command exited 0 without ncu).
OUT=gpurun_out/r2; mkdir -p $OUT
NCU="ncu --set full --clock-control none --import-source on"
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-epoch --no-extra"
$B > $OUT/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_r2.csv $B > $OUT/ncu_bench.log 2>&1
echo "launch list rc=$?"
P="python tools/profile_step.py --workload reddit --k 32"
$P > $OUT/plain_p1.log 2>&1 && $NCU -k regex:'spgemm_fwd_banked|sspmm_bwd|cbsr_bank|topk' -c 10 -o $OUT/prof_r2_reddit_k32 $P > $OUT/ncu_p1.log 2>&1
echo "reddit k32 rc=$?"; cat $OUT/plain_p1.log
P="python tools/profile_step.py --workload reddit --k 8 --warmup 0"
$P > $OUT/plain_p2.log 2>&1 && $NCU -k regex:'spgemm_fwd_banked|sspmm_bwd' -c 2 -o $OUT/prof_r2_reddit_k8 $P > $OUT/ncu_p2.log 2>&1
echo "reddit k8 rc=$?"; cat $OUT/plain_p2.log
P="python tools/profile_step.py --workload ogbn-products --k 32 --warmup 0"
$P > $OUT/plain_p3.log 2>&1 && $NCU -k regex:'spgemm_fwd|sspmm_bwd' -c 2 -o $OUT/prof_r2_products_k32 $P > $OUT/ncu_p3.log 2>&1
echo "products rc=$?"; cat $OUT/plain_p3.log
ls -la $OUT/*.ncu-rep

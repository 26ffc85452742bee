#!/bin/bash
# round 2, call 35: which commit moved the k = 16 forward from 2.14-2.22 ms to 2.39 ms?  The library and host code
# of five commits (git worktrees under scratch/wt, built in the container) timed with their own variant_bench.
OUT=$PWD/gpurun_out/r2; mkdir -p $OUT; : > $OUT/k16_bisect.log
for c in 07f5921 3f84c81 31b1536 ff082e8 ddb227d; do
  (cd scratch/wt/$c && timeout 300 python tools/variant_bench.py --ks 8,16,32 --tag $c >> $OUT/k16_bisect.log 2>&1)
done
timeout 300 python tools/variant_bench.py --ks 8,16,32 --tag HEAD >> $OUT/k16_bisect.log 2>&1
grep -v Warning $OUT/k16_bisect.log | cut -c1-220

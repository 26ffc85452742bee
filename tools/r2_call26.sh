#!/bin/bash
# round 2, GPU call 26 (1 GPU): the banked forward forced on short-record shapes (products, Flickr, Yelp).
OUT=gpurun_out/r2; mkdir -p $OUT
{
for w in ogbn-products flickr; do
  timeout 300 python tools/variant_bench.py --workload $w --ks 32 --tag plain_fwd
  MAXK_BANKED_MIN_RECORD=0 timeout 300 python tools/variant_bench.py --workload $w --ks 32 --tag banked_forced
done
timeout 300 python tools/variant_bench.py --workload ogbn-products --ks 16,64 --tag plain_fwd
MAXK_BANKED_MIN_RECORD=0 timeout 300 python tools/variant_bench.py --workload ogbn-products --ks 16,64 --tag banked_forced
} > $OUT/banked_short_records.log 2>&1
cat $OUT/banked_short_records.log

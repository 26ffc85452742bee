#!/bin/bash
# round 2, GPU call 1 (1 GPU): parity suite incl. the full-size oracle samples, the three orderings of
# the forward accumulation (MK_SYNC_MODE 0/1/2), per-rank kernel times of an 8-way shard vs max_nz, bench.
OUT=gpurun_out/r2; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest1.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest1.log
tail -3 $OUT/pytest1.log
for m in 0 1 2; do
  MAXK_LIB=$PWD/spgemm_gnn_b200/libmaxk_sync$m.so timeout 300 python tools/variant_bench.py --ks 8,16,32,64 --tag sync$m
done > $OUT/sync_modes.log 2>&1
cat $OUT/sync_modes.log
timeout 300 python tools/variant_bench.py --shard 8 --ks 32 --max-nz 128,256,512,1024 > $OUT/shard8_maxnz.log 2>&1
timeout 300 python tools/variant_bench.py --workload ogbn-products --ks 32 --max-nz 256,1024 >> $OUT/shard8_maxnz.log 2>&1
timeout 300 python tools/variant_bench.py --workload ogbn-products --shard 8 --ks 32 --max-nz 256,1024 >> $OUT/shard8_maxnz.log 2>&1
cat $OUT/shard8_maxnz.log
timeout 600 python bench.py > $OUT/bench1.json 2> $OUT/bench1.err; echo "bench rc=$?"
tail -c 1500 $OUT/bench1.err; python -c "
import json;d=json.loads(open('$OUT/bench1.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}); print(d['roofline']); print(d['e2e']); print(d['cpu_baseline']); print(d['products']); print(d['flickr'])
for r in d['ksweep']['rows']: print(r)
print(d['ksweep']['cusparse_dense_spmm']); print(d['sage_epoch']); print(d['kernels'])"

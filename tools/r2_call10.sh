#!/bin/bash
# round 2, GPU call 10 (1 GPU): parity suite with the phased forward, virtual-rank peer check (phases in
# rounds 1 and 3), top-k timing, the cp.async edge-staging variant of the backward (MK_EDGE_CPASYNC).
OUT=gpurun_out/r2; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest10.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest10.log
timeout 600 python tools/peer_check.py virtual 4 --all-k > $OUT/peer_virtual10.log 2>&1; echo "virtual rc=$?"; grep -c OK $OUT/peer_virtual10.log; tail -3 $OUT/peer_virtual10.log
{
for w in reddit ogbn-products flickr; do
  timeout 300 python tools/variant_bench.py --workload $w --ks 32 --topk --tag product
  MAXK_LIB=$PWD/spgemm_gnn_b200/libmaxk_cpasync.so timeout 300 python tools/variant_bench.py --workload $w --ks 32 --tag cpasync_bwd
done
timeout 300 python tools/variant_bench.py --workload reddit --ks 8,16,64 --topk --tag product
MAXK_LIB=$PWD/spgemm_gnn_b200/libmaxk_cpasync.so timeout 300 python tools/variant_bench.py --workload reddit --ks 8,16,64 --tag cpasync_bwd
} > $OUT/cpasync_edges.log 2>&1
cat $OUT/cpasync_edges.log

#!/bin/bash
# round 2, GPU call 12 (1 GPU): LayerNorm epilogue inside the forward SpGEMM (parity + epoch time),
# the tile top-k kernel with the bitwise search (MK_TILE_BITWISE) against the interpolation search.
OUT=gpurun_out/r2; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest12.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest12.log
{
MAXK_TOPK_TILE=1 timeout 300 python tools/variant_bench.py --workload reddit --ks 32 --topk --tag tile_interp
MAXK_TOPK_TILE=1 MAXK_LIB=$PWD/spgemm_gnn_b200/libmaxk_tilebw.so timeout 300 python tools/variant_bench.py --workload reddit --ks 8,32,64 --topk --tag tile_bitwise
} > $OUT/topk_tile_bitwise.log 2>&1
cat $OUT/topk_tile_bitwise.log
for f in 1 0; do
  MAXK_FUSED_LN=$f timeout 600 python tools/epoch_profile.py --model maxk-sage --tf32 > $OUT/epoch_ln$f.txt 2>&1; echo "epoch_profile fused_ln=$f rc=$?"; grep "epoch ms" $OUT/epoch_ln$f.txt
done

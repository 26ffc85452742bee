#!/bin/bash
# First GPU calls of the next round: what round 1 built but could not measure for lack of GPU time.
#   1 GPU : tools/round2_first_call.sh one     (epoch with / without MAXK_ALIGN_GEMM, ncu of the peer kernels in virtual mode)
#   8 GPUs: tools/round2_first_call.sh eight   (peer exchange stress + NCCL-vs-peer scaling of bench.py)
OUT=gpurun_out/round2; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
case ${1:-one} in
one)
  for a in 0 1; do
    MAXK_ALIGN_GEMM=$a python -m spgemm_gnn_b200.train --dataset reddit --model sage --maxk 32 --epochs 12 --norm --cuda_graph \
      > $OUT/epoch_align$a.log 2>&1; tail -1 $OUT/epoch_align$a.log
  done
  CUDA_DEVICE_MAX_CONNECTIONS=32 python tools/peer_check.py virtual 4 --all-k > $OUT/peer_virtual_allk.log 2>&1; tail -1 $OUT/peer_virtual_allk.log
  CUDA_DEVICE_MAX_CONNECTIONS=32 python tools/peer_check.py virtual 4 --all-k --push-mode 2 > $OUT/peer_virtual_mode2.log 2>&1; tail -1 $OUT/peer_virtual_mode2.log
  CUDA_DEVICE_MAX_CONNECTIONS=32 ncu --set full --clock-control none --import-source on -k regex:'peer_|cbsr_bank_kernel' \
    -o $OUT/peer_virtual python tools/peer_check.py virtual 4 > $OUT/ncu_peer_virtual.log 2>&1
  ;;
eight)
  $TR --nproc-per-node 8 --master-port 29661 tools/peer_check.py dist --bench --products --stress > $OUT/peer_stress8.log 2>&1
  grep -v '^\*\|OMP_NUM' $OUT/peer_stress8.log | tail -8
  $TR --nproc-per-node 8 --master-port 29662 tools/peer_check.py dist --bench --products --stress --push-mode 2 > $OUT/peer_stress8_mode2.log 2>&1
  grep -v '^\*\|OMP_NUM' $OUT/peer_stress8_mode2.log | tail -8
  for p in 1 0; do for n in 2 4 8; do
    MAXK_PEER_EXCHANGE=$p $TR --nproc-per-node $n --master-port $((29670 + n)) bench.py --gpus $n --steps 30 --warmup 5 \
      > $OUT/bench_peer$p.$n.log 2>&1
    tail -1 $OUT/bench_peer$p.$n.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('peer=$p gpus', d['n_gpus'], 'ms/layer %.3f' % d['ms_per_step'], 'epoch %.2f ms' % d['sage_epoch']['ms_per_epoch'])"
  done; done
  ;;
esac

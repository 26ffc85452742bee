#!/bin/bash
# round 2, GPU call 9 (8 GPUs): the all-gather fused into the forward SpGEMM (pusher CTAs) against the
# NCCL form and the copy-engine form, forward and backward timed on their own, pusher-count sweep;
# then bench.py at 8 ranks with the defaults.
OUT=gpurun_out/r2; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
export MAXK_PEER_TIMEOUT_MS=15000
timeout 420 $TR --nproc-per-node 8 --master-port 29691 tools/peer_check.py dist --bench --products --sweep 2>&1 \
  | grep -v '^\*\|OMP_NUM\|^W1\|^$\|NCCL version' > $OUT/peer_sweep8.log; echo "sweep rc=${PIPESTATUS[0]}"
cat $OUT/peer_sweep8.log
timeout 420 $TR --nproc-per-node 8 --master-port 29692 bench.py --gpus 8 --steps 20 --warmup 5 > $OUT/bench_n8b.json 2> $OUT/bench_n8b.err; echo "bench rc=$?"
tail -c 300 $OUT/bench_n8b.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2/bench_n8b.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('n_gpus','value','ms_per_step','gpu_launches')}); print('parity',d['parity']); print('products',d['products']); print('epoch',d['sage_epoch']['ms_per_epoch']); print('kernels',{k:d['kernels'][k] for k in ('spgemm_fwd_ms','sspmm_bwd_ms')}); print('e2e', d['e2e']['ms_per_step'])
PY

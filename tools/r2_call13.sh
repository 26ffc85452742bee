#!/bin/bash
# round 2, GPU call 13 (N GPUs, N = $1): the forward in source-block phases against the one-launch forward,
# for every way of moving the rows (NCCL, pusher CTAs, copy engines); then bench.py at N ranks.
N=${1:-8}
OUT=gpurun_out/r2; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
export MAXK_PEER_TIMEOUT_MS=15000
timeout 420 $TR --nproc-per-node $N --master-port 29693 tools/peer_check.py dist --bench --products --sweep 2>&1 \
  | grep -v '^\*\|OMP_NUM\|^W1\|^$\|NCCL version' > $OUT/peer_phases$N.log; echo "sweep rc=${PIPESTATUS[0]}"
cat $OUT/peer_phases$N.log
timeout 420 $TR --nproc-per-node $N --master-port 29694 bench.py --gpus $N --steps 20 --warmup 5 > $OUT/bench_n${N}c.json 2> $OUT/bench_n${N}c.err; echo "bench rc=$?"
tail -c 300 $OUT/bench_n${N}c.err
python - <<PY
import json
d=json.loads(open('$OUT/bench_n${N}c.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('n_gpus','value','ms_per_step','gpu_launches')}); print('parity',d['parity']); print('products',d['products']); print('epoch',d['sage_epoch']['ms_per_epoch']); print('kernels',{k:d['kernels'][k] for k in ('spgemm_fwd_ms','sspmm_bwd_ms')}); print('e2e', d['e2e']['ms_per_step'])
PY

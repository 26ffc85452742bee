#!/bin/bash
# round 2, GPU call 19 (N GPUs): NVLink multicast forms of the two exchanges (symmetric-memory windows,
# mk_peer_push_mc / mk_peer_reduce_scatter_mc) against NCCL and the unicast peer forms; then bench.py.
N=${1:-2}
OUT=gpurun_out/r2; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
export MAXK_PEER_TIMEOUT_MS=15000
timeout 420 $TR --nproc-per-node $N --master-port 29711 tools/peer_check.py dist --bench --products --sweep 2>&1 \
  | grep -v '^\*\|OMP_NUM\|^W1\|^$\|NCCL version' > $OUT/peer_mc$N.log; echo "sweep rc=${PIPESTATUS[0]}"
cat $OUT/peer_mc$N.log
timeout 420 $TR --nproc-per-node $N --master-port 29712 tools/peer_check.py dist --stress 2>&1 \
  | grep -v '^\*\|OMP_NUM\|^W1\|^$\|NCCL version' > $OUT/peer_mc_stress$N.log; echo "stress rc=${PIPESTATUS[0]}"; tail -4 $OUT/peer_mc_stress$N.log
timeout 420 $TR --nproc-per-node $N --master-port 29713 bench.py --gpus $N --steps 20 --warmup 5 > $OUT/bench_n${N}d.json 2> $OUT/bench_n${N}d.err; echo "bench rc=$?"
tail -c 300 $OUT/bench_n${N}d.err
python - <<PY
import json
d=json.loads(open('$OUT/bench_n${N}d.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('n_gpus','value','ms_per_step','gpu_launches')}); print('parity',d['parity']); print('products',d['products']); print('epoch',d['sage_epoch']['ms_per_epoch']); print('kernels',{k:d['kernels'][k] for k in ('spgemm_fwd_ms','sspmm_bwd_ms')}); print('e2e', d['e2e']['ms_per_step'])
PY

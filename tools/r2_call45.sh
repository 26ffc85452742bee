#!/bin/bash
# round 2, call 45 (8 GPUs): bench.py --gpus 8 with the final code of the round (parity record, products sub-record, epoch).
OUT=gpurun_out/r2; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
export MAXK_PEER_TIMEOUT_MS=20000
timeout 420 $TR --nproc-per-node 8 --master-port 29745 bench.py --gpus 8 --steps 20 --warmup 5 > $OUT/bench_n8_call45.json 2> $OUT/bench_n8_call45.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2/bench_n8_call45.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus','gpu_launches')}); print({k:d['parity'][k] for k in ('ok','fwd_max_rel','bwd_max_rel','fwd_bit_equal_peer_vs_nccl')}); print(d['parity']['exchange']['forward'], d['parity']['exchange']['multicast'])
print({k:d['products'][k] for k in ('ms_per_layer','fwd_ms','bwd_ms')}); print(d['sage_epoch']['ms_per_epoch']); print(d['e2e']['ms_per_step']); print(d['kernels'].get('spgemm_fwd_ms'), d['kernels'].get('sspmm_bwd_ms'))
PY
tail -c 400 $OUT/bench_n8_call45.err

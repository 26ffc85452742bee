#!/bin/bash
# round 2, GPU call 5 (8 GPUs): bench.py at 8 ranks (parity record, products record, epoch), then the
# peer form against the NCCL form on the Reddit and products shapes.
OUT=gpurun_out/r2; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 8 --master-port 29672 bench.py --gpus 8 --steps 20 --warmup 5 > $OUT/bench_n8.json 2> $OUT/bench_n8.err; echo "bench rc=$?"
tail -c 400 $OUT/bench_n8.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2/bench_n8.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('n_gpus','value','ms_per_step','gpu_launches')}); print('parity',d['parity']); print('products',d['products']); print('epoch',d['sage_epoch']['ms_per_epoch']); print('kernels',{k:d['kernels'][k] for k in ('spgemm_fwd_ms','sspmm_bwd_ms')}); print('e2e', d['e2e']['ms_per_step'])
PY
timeout 600 $TR --nproc-per-node 8 --master-port 29671 tools/peer_check.py dist --bench --products > $OUT/peer_dist8.log 2>&1; echo "peer_check rc=$?"
grep -v '^\*\|OMP_NUM\|^W1' $OUT/peer_dist8.log | tail -8

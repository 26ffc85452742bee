#!/bin/bash
# round 2, GPU call 4 (2 GPUs): the overlapped exchange for real -- peer form against the NCCL form
# (forward bit-equal, 200 back-to-back repetitions), timings, and bench.py with its parity record.
OUT=gpurun_out/r2; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 2 --master-port 29661 tools/peer_check.py dist --bench --stress > $OUT/peer_dist2.log 2>&1; echo "peer_check rc=$?"
grep -v '^\*\|OMP_NUM\|^W1' $OUT/peer_dist2.log | tail -12
timeout 600 $TR --nproc-per-node 2 --master-port 29662 bench.py --gpus 2 --steps 20 --warmup 5 > $OUT/bench_n2.json 2> $OUT/bench_n2.err; echo "bench rc=$?"
tail -c 600 $OUT/bench_n2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2/bench_n2.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('n_gpus','value','ms_per_step','gpu_launches')}); print('parity',d['parity']); print('products',d['products']); print('epoch',d['sage_epoch']); print('kernels',{k:d['kernels'][k] for k in ('spgemm_fwd_ms','sspmm_bwd_ms')}); print('e2e', d['e2e']['ms_per_step'])
PY

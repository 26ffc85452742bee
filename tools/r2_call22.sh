#!/bin/bash
# round 2, GPU call 22 (N GPUs): bench.py at N ranks with the final defaults.
N=${1:-4}
OUT=gpurun_out/r2; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node $N --master-port 29731 bench.py --gpus $N --steps 20 --warmup 5 > $OUT/bench_n${N}e.json 2> $OUT/bench_n${N}e.err; echo "bench rc=$?"
tail -c 300 $OUT/bench_n${N}e.err
python - <<PY
import json
d=json.loads(open('$OUT/bench_n${N}e.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('n_gpus','value','ms_per_step','gpu_launches')}); print('parity',d['parity']); print('products',d['products']); print('epoch',d['sage_epoch']['ms_per_epoch']); print('kernels',{k:d['kernels'][k] for k in ('spgemm_fwd_ms','sspmm_bwd_ms')}); print('e2e', d['e2e']['ms_per_step'])
PY

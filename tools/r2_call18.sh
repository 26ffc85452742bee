#!/bin/bash
# round 2, GPU call 18 (2 GPUs): bench.py at 2 ranks (new parity record: peer form, NCCL form, one-GPU rows),
# the 2-GPU peer test, and a probe of CUDA multicast memory through torch's symmetric-memory allocator.
OUT=gpurun_out/r2; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
export MAXK_PEER_TIMEOUT_MS=20000
timeout 300 $TR --nproc-per-node 2 --master-port 29701 tools/symm_probe.py 2>&1 | grep -v '^\*\|OMP_NUM\|^W1\|^$' > $OUT/symm_probe2.log; echo "probe rc=${PIPESTATUS[0]}"; cat $OUT/symm_probe2.log
timeout 600 $TR --nproc-per-node 2 --master-port 29702 bench.py --gpus 2 --steps 20 --warmup 5 > $OUT/bench_n2b.json 2> $OUT/bench_n2b.err; echo "bench rc=$?"
tail -c 400 $OUT/bench_n2b.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2/bench_n2b.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('n_gpus','value','ms_per_step','gpu_launches')}); print('parity',d['parity']); print('products',d['products']); print('epoch',d['sage_epoch']['ms_per_epoch']); print('kernels',{k:d['kernels'][k] for k in ('spgemm_fwd_ms','sspmm_bwd_ms')}); print('e2e', d['e2e']['ms_per_step'])
PY
timeout 600 python -m pytest tests/test_gpu_peer.py -m gpu -x -q > $OUT/pytest_peer2.log 2>&1; echo "peer pytest rc=$?"; tail -3 $OUT/pytest_peer2.log

#!/bin/bash
# round 2, call 49 (2 GPUs): the multi-GPU tests and the 2-GPU bench line after the forward's per-width instantiations.
OUT=gpurun_out/r2; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
export MAXK_PEER_TIMEOUT_MS=20000
timeout 900 python -m pytest tests/test_gpu_peer.py tests/test_gpu_binding.py -x -q -m gpu > $OUT/pytest49.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest49.log
timeout 420 $TR --nproc-per-node 2 --master-port 29752 tools/peer_check.py dist --bench 2>&1 \
  | grep -v '^\*\|OMP_NUM\|^W1\|^$\|NCCL version' > $OUT/peer_dist2_call49.log; echo "peer_check rc=${PIPESTATUS[0]}"
tail -12 $OUT/peer_dist2_call49.log
timeout 600 $TR --nproc-per-node 2 --master-port 29753 bench.py --gpus 2 --steps 20 --warmup 5 > $OUT/bench_n2_call49.json 2> $OUT/bench_n2_call49.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2/bench_n2_call49.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus','gpu_launches')}); print(d['parity']); print(d['products']); print(d['sage_epoch']); print(d['e2e'])
PY
timeout 300 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > $OUT/bench_ref_call49.json 2> $OUT/bench_ref_call49.err; echo "reference arm rc=$?"; tail -c 700 $OUT/bench_ref_call49.json

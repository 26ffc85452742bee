#!/bin/bash
# bench.py at 1/2/4/8 GPUs on one box (4-, 2-, 1-GPU runs side by side on disjoint GPUs).
W=${1:-reddit}; OUT=${2:-gpurun_out/scaling}; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
CUDA_VISIBLE_DEVICES=0,1,2,3 $TR --nproc-per-node 4 --master-port 29601 bench.py --gpus 4 --workload $W --steps 30 --warmup 5 > $OUT/bench_$W.4.log 2>&1 &
CUDA_VISIBLE_DEVICES=4,5 $TR --nproc-per-node 2 --master-port 29602 bench.py --gpus 2 --workload $W --steps 30 --warmup 5 > $OUT/bench_$W.2.log 2>&1 &
CUDA_VISIBLE_DEVICES=6 python bench.py --workload $W --steps 30 --warmup 5 --no-cpu-baseline > $OUT/bench_$W.1.log 2>&1 &
wait
$TR --nproc-per-node 8 --master-port 29603 bench.py --gpus 8 --workload $W --steps 30 --warmup 5 > $OUT/bench_$W.8.log 2>&1
for n in 1 2 4 8; do
  tail -1 $OUT/bench_$W.$n.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bench $W gpus', d['n_gpus'], 'ms/step %.3f' % d['ms_per_step'], 'edges/s %.3e' % d['value'], 'fwd %.3f bwd %.3f' % (d['kernels']['spgemm_fwd_ms'], d['kernels']['sspmm_bwd_ms']), 'e2e %.2f ms' % d['e2e']['ms_per_step'], 'sage epoch %.2f ms' % d['sage_epoch']['ms_per_epoch'])" 2>&1 | tail -1
done

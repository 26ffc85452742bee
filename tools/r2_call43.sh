#!/bin/bash
# round 2, call 43 (1 GPU): final validation -- full GPU suite, smoke(), default bench.py, then the ncu evidence of the
# final code: launch list of bench.py and one --set full capture of the Reddit k = 32 / k = 16 kernels.
OUT=gpurun_out/r2; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q -s > $OUT/pytest43.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|flickr shape|loss " $OUT/pytest43.log | tail -8
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke43.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke43.log
timeout 900 python bench.py > $OUT/bench43.json 2> $OUT/bench43.err; echo "bench rc=$?"; tail -c 300 $OUT/bench43.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2/bench43.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}); print(d['roofline']['frac'], d['e2e']['ms_per_step'], d['cpu_baseline']['value'])
for r in d['ksweep']['rows']: print({k:(round(v,4) if isinstance(v,float) else v) for k,v in r.items() if k in ('k','fwd_ms','bwd_ms','topk_ms','fwd_frac_of_peak','bwd_frac_of_peak')})
print(d['sage_epoch']['ms_per_epoch'], d['kernels']['maxk_topk_cbsr_ms'], d['products']['fwd_ms'], d['products']['bwd_ms'])
PY
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-epoch --no-extra"
$B > $OUT/plain_bench43.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_r2_final.csv $B > $OUT/ncu_bench43.log 2>&1
echo "launch list rc=$?"
P="python tools/profile_step.py --workload reddit --k 32"
$P > $OUT/plain_p43.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'spgemm_fwd_banked|sspmm_bwd|cbsr_bank|topk' -c 10 -f -o $OUT/prof_r2_final_reddit_k32 $P > $OUT/ncu_p43.log 2>&1
echo "reddit k32 capture rc=$?"
ls -la $OUT/*final*

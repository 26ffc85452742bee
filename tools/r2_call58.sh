#!/bin/bash
# round 2, GPU call 58 (1 GPU): full GPU suite (incl. the Flickr-shape loss curve), smoke(), default bench.py.
OUT=gpurun_out/r2; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q -s > $OUT/pytest58.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|flickr shape|loss " $OUT/pytest58.log | tail -12
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke58.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke58.log
timeout 900 python bench.py > $OUT/bench58.json 2> $OUT/bench58.err; echo "bench rc=$?"; tail -c 600 $OUT/bench58.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2/bench58.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}); print(d['roofline']); print(d['e2e']); print(d['cpu_baseline']); print(d['products']); print(d['flickr'])
for r in d['ksweep']['rows']: print(r)
print(d['sage_epoch']); print(d['kernels'])
PY

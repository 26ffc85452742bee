#!/bin/bash
OUT=gpurun_out/r2; mkdir -p $OUT
timeout 600 python tools/peer_check.py virtual 4 --all-k > $OUT/peer_virtual8.log 2>&1; echo "virtual rc=$?"; grep -c OK $OUT/peer_virtual8.log; tail -4 $OUT/peer_virtual8.log
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest8.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest8.log

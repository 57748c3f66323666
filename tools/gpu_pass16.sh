#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 70 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 10 python tools/diag_callback.py > $OUT/r01p_racecheck_all.log 2>&1
tail -15 $OUT/r01p_racecheck_all.log | cut -c1-220

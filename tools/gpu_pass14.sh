#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 200 python tools/diag_callback.py > $OUT/r01m_diag_callback.log 2>&1; cat $OUT/r01m_diag_callback.log | cut -c1-200

#!/bin/bash
# round 2, pass zg: the nd tier and the C++ API device test on the final header
set -u
OUT=gpurun_out; mkdir -p $OUT
( time timeout 330 python -m pytest tests/test_gpu_z_cpp_api.py tests/test_gpu_nd.py -m gpu -x -q -k "not soak and not cluster" ) > $OUT/r02zg_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/r02zg_pytest.log
tail -7 $OUT/r02zg_pytest.log

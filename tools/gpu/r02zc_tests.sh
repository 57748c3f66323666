#!/bin/bash
# round 2, pass zc: the part of the GPU tier that pass zb did not reach (it stopped at a test shape too small to fuse)
set -u
OUT=gpurun_out; mkdir -p $OUT
( time timeout 1500 python -m pytest tests/test_gpu_nd.py tests/test_gpu_r2c.py tests/test_gpu_z_cpp_api.py -m gpu -x -q -k "not (r2c_c2r_nd or fused_equals or chain or cluster or rejects or persistent_tile)" ) > $OUT/r02zc_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/r02zc_pytest.log
tail -8 $OUT/r02zc_pytest.log

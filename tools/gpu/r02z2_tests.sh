#!/bin/bash
# round 2, pass z2: the full GPU tier once more (the pass-z run stopped at a wrong assertion in a new test)
set -u
OUT=gpurun_out; mkdir -p $OUT
( time timeout 1700 python -m pytest tests -m gpu -x -q ) > $OUT/r02z2_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/r02z2_pytest.log
tail -8 $OUT/r02z2_pytest.log

#!/bin/bash
# round 2, pass i: the driver's GPU tier, as the driver runs it, then smoke()
set -u
OUT=gpurun_out; mkdir -p $OUT
( time timeout 3300 python -m pytest tests/ -x -q -m gpu ) > $OUT/r02i_pytest.log 2>&1
echo "pytest rc=$?" >> $OUT/r02i_pytest.log
tail -15 $OUT/r02i_pytest.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2

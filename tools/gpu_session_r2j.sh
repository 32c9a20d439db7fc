#!/bin/bash
# session J: phase fusion -- parity tests, then A/B timing (UAPIC_FUSE_BA, UAPIC_FUSE_SORT)
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_onepass.py -x -q -m gpu > $O/r2j_tests.log 2>&1; echo "tests rc=$?" | tee -a $O/r2j_tests.log; tail -15 $O/r2j_tests.log
: > $O/r2j_ab.log
for cfg in "0 1" "1 1" "1 2" "1 4" "1 8" "1 0"; do
  set -- $cfg
  for np in 2000000 12500000; do
    echo "FUSE=$1 SORT=$2" >> $O/r2j_ab.log
    UAPIC_FUSE_BA=$1 UAPIC_FUSE_SORT=$2 timeout 300 python tools/time_phases.py $np lean 2>&1 | grep -E "whole step|np=" >> $O/r2j_ab.log
  done
done
cat $O/r2j_ab.log

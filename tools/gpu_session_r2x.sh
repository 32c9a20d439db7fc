#!/bin/bash
# the shipped default (two tau samples per lane) once more: efd parity tests and the reference-size timing
O=gpurun_out; mkdir -p $O
timeout 120 python -m pytest tests/test_gpu_efd.py -q -x 2>&1 | tail -3 | tee $O/r2x_efd_tests.log
timeout 60 python tools/bench_efd.py > $O/r2x_bench_204800.json 2>/dev/null; tail -1 $O/r2x_bench_204800.json | cut -c1-420

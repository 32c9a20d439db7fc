#!/bin/bash
# external-field program: parity tests, then timing at the reference's size and at 1e7 particles
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_efd.py -q -x 2>&1 | tail -15 | tee gpurun_out/r2efd_tests.log
timeout 300 python tools/bench_efd.py > gpurun_out/r2efd_bench_204800.json 2> gpurun_out/r2efd_bench.err; tail -1 gpurun_out/r2efd_bench_204800.json
timeout 300 python tools/bench_efd.py --particles 10000000 --cpu-particles 400000 > gpurun_out/r2efd_bench_1e7.json 2>> gpurun_out/r2efd_bench.err; tail -1 gpurun_out/r2efd_bench_1e7.json
tail -3 gpurun_out/r2efd_bench.err

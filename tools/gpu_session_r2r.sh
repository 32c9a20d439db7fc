#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
UAPIC_B200_LIB=$PWD/uapic.jl_b200/variants/libuapic_b200_tmem.so timeout 600 python -m pytest tests/test_gpu_onepass.py tests/test_gpu_referee.py -q -x 2>&1 | tail -4
bash tools/gpu_ab.sh r2r 2000000,12500000 default tmem default tmem
UAPIC_B200_LIB=$PWD/uapic.jl_b200/variants/libuapic_b200_tmem.so timeout 300 ncu --clock-control none --metrics gpu__time_duration.sum,l1tex__data_pipe_lsu_wavefronts.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active -k regex:k_onepass_a -s 3 -c 1 --csv --log-file $O/r2r_tmem_ncu.csv python tools/time_phases.py 2000000 lean > /dev/null 2>&1
grep -E "k_onepass_a" $O/r2r_tmem_ncu.csv | cut -d, -f13-15

#!/bin/bash
# A/B of the external-field kernel's twiddle source: registers (default build) vs shared-memory table (variant)
O=gpurun_out; mkdir -p $O
for rep in 1 2; do
  python tools/bench_efd.py --particles 2000000 --cpu-particles 1000 --reps 10 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('regs  ', d['gpu_kernel'])"
  UAPIC_B200_LIB=uapic.jl_b200/variants/libuapic_b200_efdlds.so python tools/bench_efd.py --particles 2000000 --cpu-particles 1000 --reps 10 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('shared', d['gpu_kernel'])"
done | tee $O/r2k2_efd_twiddle_ab.log
timeout 600 python -m pytest tests/test_gpu_efd.py -q -x 2>&1 | tail -3 | tee -a $O/r2k2_efd_twiddle_ab.log

#!/bin/bash
# two tau samples per lane (LaneTau2, UAPIC_EFD_SPL2=1) against one (default): parity of the new policy, then A/B timing
O=gpurun_out; mkdir -p $O
UAPIC_EFD_SPL2=1 timeout 200 python -m pytest tests/test_gpu_efd.py -q -x 2>&1 | tail -4 | tee $O/r2w_efd_spl2.log
for v in 0 1; do
  UAPIC_EFD_SPL2=$v timeout 100 python tools/bench_efd.py --particles 2000000 --cpu-particles 1000 --reps 10 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('UAPIC_EFD_SPL2=$v', d['gpu_kernel'])"
done | tee -a $O/r2w_efd_spl2.log

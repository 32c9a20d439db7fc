#!/bin/bash
# end of round 2: the whole GPU suite on the final tree, smoke, and one ncu --set full capture of the external-field kernel
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5 | tee $O/r2g_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee $O/r2g_smoke.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_efd -s 3 -c 1 -o $O/r2g_efd python tools/bench_efd.py --particles 2000000 --cpu-particles 1000 --reps 1 > $O/r2g_ncu.log 2>&1; tail -2 $O/r2g_ncu.log
ncu -i $O/r2g_efd.ncu-rep --page raw --csv > $O/r2g_efd_ncu_full.csv 2>/dev/null; wc -c $O/r2g_efd_ncu_full.csv

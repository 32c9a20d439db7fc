#!/bin/bash
# ncu --set full of the slimmed external-field kernel (same command as r2g, for a before/after of the instruction-fetch stalls)
O=gpurun_out; mkdir -p $O
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_efd -s 3 -c 1 -o $O/r2h_efd python tools/bench_efd.py --particles 2000000 --cpu-particles 1000 --reps 1 > $O/r2h_ncu.log 2>&1; tail -2 $O/r2h_ncu.log
ncu -i $O/r2h_efd.ncu-rep --page raw --csv > $O/r2h_efd_ncu_full.csv 2>/dev/null; wc -c $O/r2h_efd_ncu_full.csv

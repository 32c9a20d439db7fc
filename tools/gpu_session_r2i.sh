#!/bin/bash
# session I: wavefronts of the shipped kernels with and without their gather loads (what do 36 taps cost IN the kernels?)
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
M=gpu__time_duration.sum,l1tex__data_pipe_lsu_wavefronts.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,l1tex__t_sector_hit_rate.pct,smsp__issue_active.avg.pct_of_peak_sustained_active
for lib in default nogather; do
  if [ $lib = default ]; then unset UAPIC_B200_LIB; else export UAPIC_B200_LIB=$PWD/uapic.jl_b200/variants/libuapic_b200_$lib.so; fi
  timeout 300 ncu --clock-control none --metrics $M -k regex:k_onepass -s 6 -c 2 --csv --log-file $O/r2i_$lib.csv python tools/time_phases.py 2000000 lean > $O/r2i_$lib.log 2>&1
  python - $O/r2i_$lib.csv $lib <<'PY'
import csv,sys
lines=open(sys.argv[1]).read().splitlines()
i=[k for k,l in enumerate(lines) if l.startswith('"ID"')][0]
by={}
for r in csv.DictReader(lines[i:]): by.setdefault((r['ID'],r['Kernel Name'].split('<')[0].split('::')[-1]),{})[r['Metric Name']]=r['Metric Value']
for k,v in by.items():
    g=lambda m: float(v[m].replace(',',''))
    print(sys.argv[2], k[1], 'time %.3f ms' % (g('gpu__time_duration.sum')/1e6), 'wavefronts/particle %.1f' % (g('l1tex__data_pipe_lsu_wavefronts.sum')/2e6), 'lsu %s%% fp64 %s%% issue %s%% hit %s' % (v['l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed'], v['sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active'], v['smsp__issue_active.avg.pct_of_peak_sustained_active'], v['l1tex__t_sector_hit_rate.pct']))
PY
done
unset UAPIC_B200_LIB
timeout 300 python tools/time_phases.py 2000000 lean 2>&1 | tail -1

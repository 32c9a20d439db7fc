#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
B=tools/microbench/gather_aligned_bench
: > $O/r2p_aligned.jsonl
for args in "2000000 128 128 0.1" "2000000 128 64 0.1" "2000000 256 256 0.1"; do timeout 200 $B $args | tee -a $O/r2p_aligned.jsonl; done
timeout 300 ncu --clock-control none --metrics gpu__time_duration.sum,l1tex__data_pipe_lsu_wavefronts.sum,l1tex__t_sector_hit_rate.pct,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum -k regex:k_gather --csv --log-file $O/r2p_aligned_ncu.csv $B 2000000 128 128 0.1 > /dev/null 2>&1
python - <<'PY'
import csv
lines=open('gpurun_out/r2p_aligned_ncu.csv').read().splitlines()
i=[k for k,l in enumerate(lines) if l.startswith('"ID"')][0]
by={}
for r in csv.DictReader(lines[i:]): by.setdefault((int(r['ID']),r['Kernel Name'][:24]),{})[r['Metric Name']]=r['Metric Value']
for k,v in sorted(by.items()):
    if k[0] % 6 != 5: continue
    g=lambda m: float(v[m].replace(',',''))
    print(k[1], 'time %.3f ms' % (g('gpu__time_duration.sum')/1e6), 'wavefronts/particle %.1f' % (g('l1tex__data_pipe_lsu_wavefronts.sum')/2e6), 'hit', v['l1tex__t_sector_hit_rate.pct'], 'lsu%', v['l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed'], 'tags/particle %.1f' % (g('l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum')/2e6))
PY

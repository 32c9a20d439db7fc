#!/bin/bash
# round 2, session D (one B200): gather / deposit microbenchmark (TMA-staged smem tile, smem deposit) + ncu wavefront counts
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
B=tools/microbench/gather_deposit_bench
: > $O/r2d_microbench.jsonl
for args in "2000000 128 128 0.1" "2000000 128 64 0.1" "2000000 256 256 0.1" "2000000 128 128 0.01"; do
  timeout 300 $B $args | tee -a $O/r2d_microbench.jsonl
done
timeout 600 ncu --clock-control none --metrics gpu__time_duration.sum,l1tex__data_pipe_lsu_wavefronts.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum,l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_red.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum \
  -k regex:"k_gather|k_deposit" -s 0 -c 6 --csv --log-file $O/r2d_microbench_ncu.csv $B 2000000 128 128 0.1 > $O/r2d_ncu_stdout.log 2>&1
python - <<'PY'
import csv
lines=open('gpurun_out/r2d_microbench_ncu.csv').read().splitlines()
i=[k for k,l in enumerate(lines) if l.startswith('"ID"')][0]
rows=list(csv.DictReader(lines[i:]))
by={}
for r in rows: by.setdefault((r['ID'],r['Kernel Name'][:40]),{})[r['Metric Name']]=r['Metric Value']
for k,v in by.items(): print(k, {m.split('__')[-1][:50]:x for m,x in v.items()})
PY

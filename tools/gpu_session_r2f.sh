#!/bin/bash
# round 2, session F (one B200): E-halo layout sweep of the M6 gather, time + ncu data-pipe wavefronts per layout
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
B=tools/microbench/gather_layout_bench
: > $O/r2f_layouts.jsonl
for args in "2000000 128 128 0.1" "2000000 128 64 0.1" "2000000 256 256 0.1"; do timeout 300 $B $args | tee -a $O/r2f_layouts.jsonl; done
timeout 900 ncu --clock-control none --metrics gpu__time_duration.sum,l1tex__data_pipe_lsu_wavefronts.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sector_hit_rate.pct,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,l1tex__t_set_accesses_pipe_lsu_mem_global_op_ld.sum,l1tex__t_set_conflicts_pipe_lsu_mem_global_op_ld.sum \
  -k regex:k_gather --csv --log-file $O/r2f_layouts_ncu.csv $B 2000000 128 128 0.1 > $O/r2f_ncu_stdout.log 2>&1
python - <<'PY'
import csv
lines=open('gpurun_out/r2f_layouts_ncu.csv').read().splitlines()
i=[k for k,l in enumerate(lines) if l.startswith('"ID"')][0]
by={}
for r in csv.DictReader(lines[i:]): by.setdefault(int(r['ID']),{})[r['Metric Name']]=r['Metric Value']
names=["tile_2x4","tile_4x2","tile_8x1","tile_1x8","rows_pitch1","rows_pitch2","rows_pitch3","rows_pitch5","rows_pitch6","quarter_choice"]
for k in sorted(by):
    if k%4!=3: continue          # 1 warm-up + 3 timed launches per layout: take the last
    v=by[k]; g=lambda m: float(v[m].replace(',',''))
    wf=g('l1tex__data_pipe_lsu_wavefronts.sum')/2e6
    print(f"{names[k//4]:16s} time {g('gpu__time_duration.sum')/1e6:.3f} ms  wavefronts/particle {wf:6.1f}  per tap {(wf-12)/36:5.2f}  tag requests/particle {g('l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum')/2e6:5.1f} sectors {g('l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum')/2e6:6.1f} t_out_wf {g('l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum')/2e6:6.1f} hit {v['l1tex__t_sector_hit_rate.pct']} lsu% {v['l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed']}")
PY

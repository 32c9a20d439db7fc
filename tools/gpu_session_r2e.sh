#!/bin/bash
# round 2, session E (one B200): full GPU tests, ncu captures of the shipped kernels (full set at 2e6, DRAM bytes at 12.5e6),
# the round's bench lines of every BASELINE config that fits one GPU
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -x -q -rs > $O/r2e_tests.log 2>&1; echo "tests rc=$?" | tee -a $O/r2e_tests.log; tail -5 $O/r2e_tests.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_onepass -s 6 -c 2 -o $O/r2e_onepass python tools/time_phases.py 2000000 lean > $O/r2e_ncu.log 2>&1; tail -2 $O/r2e_ncu.log
timeout 600 ncu --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:k_onepass -s 6 -c 4 --csv --log-file $O/r2e_dram_12p5M.csv python tools/time_phases.py 12500000 lean > $O/r2e_ncu2.log 2>&1
python tools/traffic_from_ncu.py $O/r2e_dram_12p5M.csv 12500000 32 > $O/r2e_traffic.json; cat $O/r2e_traffic.json; cp $O/r2e_traffic.json profiles/r2e_traffic.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2e_launches.csv python bench.py --steps 2 --warmup 3 --particles-per-gpu 2000000 --no-e2e --no-cpu-baseline > $O/r2e_launch_bench.log 2>&1
show() { python -c "
import json
d=json.loads([l for l in open('$1') if l.startswith('{')][-1]); r=d['roofline']
print('$1', 'value %.3e ms %.3f e2e %.3e A %.3f B %.3f barrier %s frac %.3f fp64 %s' % (d['value'], d['ms_per_step'], d['e2e']['value'] if d['e2e'] else 0, r['phase_a_ms'], r['phase_b_ms'], r.get('field_barrier_ms'), r['whole_step_frac_of_hbm'], (r.get('fp64') or {}).get('frac')))"; }
timeout 900 python bench.py > $O/r2e_bench_config3.json 2> $O/r2e_bench_config3.err; show $O/r2e_bench_config3.json
timeout 600 python bench.py --workload config2 --steps 50 > $O/r2e_bench_config2.json 2> $O/r2e_bench_config2.err; show $O/r2e_bench_config2.json
timeout 600 python bench.py --workload config5 > $O/r2e_bench_config5_n1.json 2> $O/r2e_bench_config5_n1.err; show $O/r2e_bench_config5_n1.json
for e in 1e-1 1e-2 1e-3 1e-4 1e-5; do
  timeout 300 python bench.py --workload config4 --eps $e --no-cpu-baseline --steps 8 > $O/r2e_bench_config4_eps$e.json 2> $O/r2e_bench_config4_eps$e.err; show $O/r2e_bench_config4_eps$e.json
done
timeout 600 python bench.py --impl reference --steps 3 > $O/r2e_bench_reference_arm.json 2> $O/r2e_bench_reference_arm.err; head -c 300 $O/r2e_bench_reference_arm.json

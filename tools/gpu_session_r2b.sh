#!/bin/bash
# round 2, GPU session B (one B200): all GPU tests incl. referee + round-2 features, small-eps table, fused vs split solve at config 2
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -x -q > $O/r2b_tests.log 2>&1; echo "tests rc=$?" | tee -a $O/r2b_tests.log
tail -15 $O/r2b_tests.log
timeout 600 python tools/small_eps_table.py > $O/r2b_small_eps_parity.json 2> $O/r2b_small_eps.log; cat $O/r2b_small_eps.log
for split in 0 1; do
  UAPIC_SPLIT_SOLVE=$split timeout 600 python bench.py --workload config2 --steps 50 --no-cpu-baseline > $O/r2b_bench_config2_split$split.json 2> $O/r2b_bench_config2_split$split.err
  python -c "import json; d=json.load(open('$O/r2b_bench_config2_split$split.json')); r=d['roofline']; print('split=$split value %.3e ms/step %.4f A+B %.4f launches %d e2e %.3e' % (d['value'], d['ms_per_step'], r['phase_a_ms']+r['phase_b_ms'], d['gpu_launches'], d['e2e']['value']))"
done
timeout 300 python tools/time_phases.py 2000000 lean >> $O/r2b_ab.log 2>&1; cat $O/r2b_ab.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/r2b_launches_config2.csv python bench.py --workload config2 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > $O/r2b_launch_bench.log 2>&1
grep -E "k_field_solve|k_sort|k_onepass|k_fold|Memset" $O/r2b_launches_config2.csv | tail -12 | cut -d, -f5,15 

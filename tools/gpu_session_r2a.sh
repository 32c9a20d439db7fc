#!/bin/bash
# round 2, GPU session A (one B200): parity tests, pair-load A/B, fp64 peak, bench lines of the configs never run, ncu capture
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
nvidia-smi -L > $O/r2a_gpus.txt; nproc >> $O/r2a_gpus.txt
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r2a_tests.log 2>&1; echo "tests rc=$?" | tee -a $O/r2a_tests.log
tail -3 $O/r2a_tests.log
for lib in default nopair nodeposit; do
  if [ $lib = default ]; then unset UAPIC_B200_LIB; else export UAPIC_B200_LIB=$PWD/uapic.jl_b200/variants/libuapic_b200_$lib.so; fi
  timeout 300 python tools/time_phases.py 2000000 lean >> $O/r2a_ab.log 2>&1
  timeout 300 python tools/time_phases.py 12500000 lean >> $O/r2a_ab.log 2>&1
done
unset UAPIC_B200_LIB
cat $O/r2a_ab.log
timeout 120 python bench.py --peaks > $O/r2a_peaks.json 2> $O/r2a_peaks.err; cat $O/r2a_peaks.json
timeout 600 python bench.py --workload config2 --steps 50 > $O/r2a_bench_config2.json 2> $O/r2a_bench_config2.err; tail -c 600 $O/r2a_bench_config2.json
timeout 600 python bench.py --workload config5 > $O/r2a_bench_config5_n1.json 2> $O/r2a_bench_config5_n1.err; tail -c 600 $O/r2a_bench_config5_n1.json
for e in 1e-1 1e-2 1e-3 1e-4 1e-5; do
  timeout 300 python bench.py --workload config4 --eps $e --no-cpu-baseline --steps 8 > $O/r2a_bench_config4_eps$e.json 2> $O/r2a_bench_config4_eps$e.err
  python -c "import json,sys; d=json.load(open('$O/r2a_bench_config4_eps$e.json')); print('eps $e', d['value'], d['e2e']['value'], d['roofline']['phase_a_ms'], d['roofline']['phase_b_ms'])"
done
timeout 900 python bench.py > $O/r2a_bench_config3.json 2> $O/r2a_bench_config3.err; tail -c 1500 $O/r2a_bench_config3.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_onepass -s 6 -c 2 -o $O/r2a_onepass python tools/time_phases.py 2000000 lean > $O/r2a_ncu.log 2>&1; tail -2 $O/r2a_ncu.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2a_launches.csv python bench.py --steps 2 --warmup 3 --particles-per-gpu 2000000 --no-e2e --no-cpu-baseline > $O/r2a_launch_bench.log 2>&1
ls -la $O | tail -30

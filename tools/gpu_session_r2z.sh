#!/bin/bash
# final session of round 2 (one B200): whole GPU suite, smoke(), the default bench line and the reference arm
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q -rs > $O/r2z_tests.log 2>&1; echo "tests rc=$?" | tee -a $O/r2z_tests.log; tail -6 $O/r2z_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > $O/r2z_bench_config3.json 2> $O/r2z_bench_config3.err; tail -c 2600 $O/r2z_bench_config3.json
timeout 600 python bench.py --impl reference --steps 3 > $O/r2z_bench_reference_arm.json 2> /dev/null; head -c 200 $O/r2z_bench_reference_arm.json
timeout 600 python bench.py --workload config2 --steps 50 --no-cpu-baseline > $O/r2z_bench_config2.json 2> /dev/null; python -c "
import json; d=json.load(open('$O/r2z_bench_config2.json')); r=d['roofline']; print('config2 %.3e %.4f ms A+B %.4f'%(d['value'],d['ms_per_step'],r['phase_a_ms']+r['phase_b_ms']))"

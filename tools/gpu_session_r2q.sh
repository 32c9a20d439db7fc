#!/bin/bash
# session Q (4 GPUs): the driver's sequence at N = 4 -- reference arm under torchrun, then the default bench; config 2 at 4 GPUs
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
N=4; P=29700
run() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((P++)) "$@"; }
run bench.py --impl reference --gpus $N --steps 3 --warmup 3 > $O/r2q_reference_arm_n4.json 2> $O/r2q_reference_arm_n4.err; echo "reference arm lines: $(wc -l < $O/r2q_reference_arm_n4.json)"; head -c 250 $O/r2q_reference_arm_n4.json; echo
run bench.py --gpus $N --steps 10 --warmup 3 > $O/r2q_bench_config3_n4.json 2> $O/r2q_bench_config3_n4.err; echo "bench lines: $(wc -l < $O/r2q_bench_config3_n4.json)"
python -c "
import json; d=json.load(open('$O/r2q_bench_config3_n4.json')); r=d['roofline']
print('config3 n4 value %.4e ms %.3f e2e %.4e (%.2f ms) A %.2f B %.2f barrier %.4f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], r['phase_a_ms'], r['phase_b_ms'], r['field_barrier_ms']))
print('cpu_baseline', d['cpu_baseline']['value'], d['cpu_baseline']['cores'])"
run bench.py --gpus $N --workload config2 --steps 50 --no-cpu-baseline > $O/r2q_bench_config2_n4.json 2> /dev/null
python -c "
import json; d=json.load(open('$O/r2q_bench_config2_n4.json')); r=d['roofline']
print('config2 n4 value %.4e ms %.4f A+B %.4f' % (d['value'], d['ms_per_step'], r['phase_a_ms']+r['phase_b_ms']))"

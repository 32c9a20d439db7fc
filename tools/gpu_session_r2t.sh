#!/bin/bash
# final 8-GPU confirmation on the final code: NCCL parity at 2/4/8 ranks (2D and 3D paths), the driver's sequence at N = 8
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_multi.py -q -rs > $O/r2t_multi_test_n8.log 2>&1; tail -3 $O/r2t_multi_test_n8.log
N=8; P=29800
run() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((P++)) "$@"; }
run bench.py --impl reference --gpus $N --steps 3 --warmup 3 > $O/r2t_reference_arm_n8.json 2> /dev/null; echo "reference arm lines: $(wc -l < $O/r2t_reference_arm_n8.json)"
run bench.py --gpus $N --steps 10 --warmup 3 > $O/r2t_bench_config3_n8.json 2> $O/r2t_bench_config3_n8.err; echo "bench lines: $(wc -l < $O/r2t_bench_config3_n8.json)"
python -c "
import json; d=json.load(open('$O/r2t_bench_config3_n8.json')); r=d['roofline']
print('config3 n8 value %.4e ms %.3f e2e %.4e (%.2f ms) A %.2f B %.2f barrier %.4f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], r['phase_a_ms'], r['phase_b_ms'], r['field_barrier_ms']))"

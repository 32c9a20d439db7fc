#!/bin/bash
# session V (8 GPUs): NCCL / callback / peer-memory parity at 2, 4, 8 ranks; config 2 at 8 GPUs with NCCL and with the fused exchange
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_multi.py -q -rs > $O/r2v_multi_test_n8.log 2>&1; tail -3 $O/r2v_multi_test_n8.log
N=8; P=30100
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((P++)) "$@"; }
for red in nccl peer; do
  run bench.py --gpus $N --no-cpu-baseline --no-e2e --workload config2 --steps 200 --reduce $red > $O/r2u_config2_n${N}_$red.json 2> /dev/null
  python -c "
import json; d=json.load(open('$O/r2u_config2_n${N}_$red.json')); r=d['roofline']
print('config2 n$N $red value %.4e ms %.4f A+B %.4f barrier %.4f' % (d['value'], d['ms_per_step'], r['phase_a_ms']+r['phase_b_ms'], r['field_barrier_ms']))"
done

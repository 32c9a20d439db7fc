#!/bin/bash
# session U (N GPUs): the exchange step three ways -- in-library NCCL, torch callback, peer memory inside the solve kernel
N=${1:-2}
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
P=29900
run() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((P++)) "$@"; }
for red in nccl peer nccl peer; do
  run bench.py --gpus $N --no-cpu-baseline --no-e2e --workload config2 --steps 200 --reduce $red > $O/r2u_config2_n${N}_$red.json 2> /dev/null
  python -c "
import json; d=json.load(open('$O/r2u_config2_n${N}_$red.json')); r=d['roofline']
print('config2 n$N $red value %.4e ms %.4f A+B %.4f barrier %.4f launches/step %.1f' % (d['value'], d['ms_per_step'], r['phase_a_ms']+r['phase_b_ms'], r['field_barrier_ms'], d['gpu_launches']/d['steps']))"
done
for red in nccl peer; do
  run bench.py --gpus $N --no-cpu-baseline --no-e2e --steps 5 --reduce $red > $O/r2u_config3_n${N}_$red.json 2> /dev/null
  python -c "
import json; d=json.load(open('$O/r2u_config3_n${N}_$red.json')); r=d['roofline']
print('config3 n$N $red value %.4e ms %.4f barrier %.4f' % (d['value'], d['ms_per_step'], r['field_barrier_ms']))"
done

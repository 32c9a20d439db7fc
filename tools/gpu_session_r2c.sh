#!/bin/bash
# round 2, multi-GPU session C (N = $1 B200s of one box): NCCL parity test through the native path, PCIe probe with all ranks
# active, bench lines of configs 3 (strong), 5 (strong + weak) and 2 at N GPUs, both reducers
N=${1:-2}
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
P=29500
run() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((P++)) "$@"; }
nvidia-smi topo -m > $O/r2c_topo_n$N.txt 2>&1; nproc >> $O/r2c_topo_n$N.txt; lscpu | grep -E "NUMA|Model name|Socket" >> $O/r2c_topo_n$N.txt
if [ -z "$SKIP_TESTS" ]; then timeout 1500 python -m pytest ${TESTS:-tests} -x -q -m gpu -rs > $O/r2c_multi_test_n$N.log 2>&1; echo "multi test rc=$?" | tee -a $O/r2c_multi_test_n$N.log; tail -4 $O/r2c_multi_test_n$N.log; fi
run tools/pcie_probe.py > $O/r2c_pcie_probe_n$N.json 2> $O/r2c_pcie_probe_n$N.err; cat $O/r2c_pcie_probe_n$N.json
show() { python -c "
import json,sys
d=json.loads([l for l in open('$1') if l.startswith('{')][-1]); r=d['roofline']
print('$1', 'value %.3e ms %.3f e2e %.3e (%.3f ms) A %.3f B %.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'] if d['e2e'] else 0, d['e2e']['ms_per_step'] if d['e2e'] else 0, r['phase_a_ms'], r['phase_b_ms']))"; }
run bench.py --gpus $N --no-cpu-baseline > $O/r2c_bench_config3_n$N.json 2> $O/r2c_bench_config3_n$N.err; show $O/r2c_bench_config3_n$N.json
run bench.py --gpus $N --no-cpu-baseline --reduce torch --no-e2e --steps 5 > $O/r2c_bench_config3_n${N}_torchreduce.json 2> /dev/null; show $O/r2c_bench_config3_n${N}_torchreduce.json
run bench.py --gpus $N --no-cpu-baseline --workload config5-strong > $O/r2c_bench_config5_strong_n$N.json 2> $O/r2c_bench_config5_strong_n$N.err; show $O/r2c_bench_config5_strong_n$N.json
run bench.py --gpus $N --no-cpu-baseline --workload config5 > $O/r2c_bench_config5_weak_n$N.json 2> $O/r2c_bench_config5_weak_n$N.err; show $O/r2c_bench_config5_weak_n$N.json
run bench.py --gpus $N --no-cpu-baseline --workload config2 --steps 50 > $O/r2c_bench_config2_n$N.json 2> $O/r2c_bench_config2_n$N.err; show $O/r2c_bench_config2_n$N.json
run bench.py --gpus $N --no-cpu-baseline --workload config2 --steps 50 --reduce torch --no-e2e > $O/r2c_bench_config2_n${N}_torchreduce.json 2> /dev/null; show $O/r2c_bench_config2_n${N}_torchreduce.json
if [ "$N" = 8 ]; then
  N=4
  run bench.py --gpus 4 --no-cpu-baseline --workload config5-strong > $O/r2c_bench_config5_strong_n4.json 2> $O/r2c_bench_config5_strong_n4.err; show $O/r2c_bench_config5_strong_n4.json
  run bench.py --gpus 4 --no-cpu-baseline --workload config5 > $O/r2c_bench_config5_weak_n4.json 2> $O/r2c_bench_config5_weak_n4.err; show $O/r2c_bench_config5_weak_n4.json
  run tools/pcie_probe.py > $O/r2c_pcie_probe_n4.json 2> $O/r2c_pcie_probe_n4.err; cat $O/r2c_pcie_probe_n4.json
fi

#!/bin/bash
# session N: 3D path (graph replay) tests + bench, compute-sanitizer over every session path incl. the round-2 kernels
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 120 python -m pytest tests/test_gpu_mrc3d.py -q 2>&1 | tail -3
timeout 100 python tools/bench_uapic3d.py --cpu-outer 0 > $O/r2m_bench_uapic3d_graph.json; cat $O/r2m_bench_uapic3d_graph.json
UAPIC3D_NO_GRAPH=1 timeout 100 python tools/bench_uapic3d.py --cpu-outer 0
: > $O/r2n_sanitizer.txt
for tool in memcheck racecheck; do
  echo "=== compute-sanitizer --tool $tool python profiles/sanitizer_driver.py" >> $O/r2n_sanitizer.txt
  UAPIC3D_NO_GRAPH=1 timeout 900 compute-sanitizer --tool $tool python profiles/sanitizer_driver.py 2>&1 | grep -vE "^$" | tail -25 >> $O/r2n_sanitizer.txt
done
cat $O/r2n_sanitizer.txt | cut -c1-200

#!/bin/bash
O=gpurun_out; mkdir -p $O; : > $O/r2s_sanitizer_efd.txt
for tool in memcheck racecheck; do
  echo "=== compute-sanitizer --tool $tool python profiles/sanitizer_driver_efd.py" >> $O/r2s_sanitizer_efd.txt
  timeout 400 compute-sanitizer --tool $tool python profiles/sanitizer_driver_efd.py 2>&1 | grep -vE "^$" | tail -12 >> $O/r2s_sanitizer_efd.txt
done
cat $O/r2s_sanitizer_efd.txt

#!/bin/bash
# memcheck + racecheck of the default (two-samples-per-lane) external-field kernel, short driver
O=gpurun_out; mkdir -p $O; : > $O/r2y_sanitizer_efd_spl2.txt
for tool in memcheck racecheck; do
  echo "=== EFD_SANITIZER_LANES_ONLY=1 compute-sanitizer --tool $tool python profiles/sanitizer_driver_efd.py" >> $O/r2y_sanitizer_efd_spl2.txt
  EFD_SANITIZER_LANES_ONLY=1 timeout 20 compute-sanitizer --tool $tool python profiles/sanitizer_driver_efd.py 2>&1 | grep -vE "^$" | tail -6 >> $O/r2y_sanitizer_efd_spl2.txt
done
cat $O/r2y_sanitizer_efd_spl2.txt

#!/bin/bash
# A/B of library variants (tools/build_variant.sh NAME ...): tools/gpu_ab.sh TAG NP[,NP] default NAME1 NAME2 ...
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
tag=$1; nps=$2; shift 2
for lib in "$@"; do
  if [ $lib = default ]; then unset UAPIC_B200_LIB; else export UAPIC_B200_LIB=$PWD/uapic.jl_b200/variants/libuapic_b200_$lib.so; fi
  for np in ${nps//,/ }; do timeout 300 python tools/time_phases.py $np lean 2>&1 | grep -v "whole step" >> $O/${tag}_ab.log; done
done
cat $O/${tag}_ab.log

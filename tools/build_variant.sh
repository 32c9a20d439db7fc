#!/bin/bash
# tools/build_variant.sh NAME -DUAPIC_OP_X=v ...  ->  uapic.jl_b200/variants/libuapic_b200_NAME.so (git-ignored, travels with gpurun)
# A/B-test it with UAPIC_B200_LIB=uapic.jl_b200/variants/libuapic_b200_NAME.so (uapic.jl_b200/_lib.py).
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
mkdir -p "$root/uapic.jl_b200/variants" "/tmp/uapic_variant_$name"
make -C "$root/uapic.jl_b200/csrc" -j4 DEFS="$*" OUT="../variants/libuapic_b200_$name.so" OBJDIR="/tmp/uapic_variant_$name"

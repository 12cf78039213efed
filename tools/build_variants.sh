#!/bin/bash
# Build flag variants of the engine for on-GPU A/B timing (development tool; outputs under variants/, git-ignored *.so)
cd "$(dirname "$0")/../climt_b200" || exit 1
mkdir -p ../variants
BASE="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --expt-relaxed-constexpr -Xcompiler -fPIC -shared"
build() { tag=$1; shift; nvcc $BASE "$@" -o ../variants/libclimt_b200_$tag.so csrc/lw_engine.cu csrc/sw_engine.cu -lcudart -ldl 2>&1 | grep -E "error" ; echo built $tag; }
build fmad_mb3 --fmad=true -DCB_UNITS_MIN_BLOCKS=3 &
build fmad_mb4 --fmad=true -DCB_UNITS_MIN_BLOCKS=4 &
build fmad_mb5 --fmad=true -DCB_UNITS_MIN_BLOCKS=5 &
build fmad_mb6 --fmad=true -DCB_UNITS_MIN_BLOCKS=6 &
wait
ls -la ../variants

#!/bin/bash
# Build flag variants of the engine for on-GPU A/B timing (development tool; outputs under variants/, git-ignored *.so)
cd "$(dirname "$0")/../climt_b200" || exit 1
mkdir -p ../variants
rm -f ../variants/*.so
BASE="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --expt-relaxed-constexpr -Xcompiler -fPIC -shared --fmad=true"
build() { tag=$1; shift; nvcc $BASE "$@" -o ../variants/libclimt_b200_$tag.so csrc/lw_engine.cu csrc/sw_engine.cu csrc/gray_engine.cu -lcudart -ldl 2>&1 | grep -E "error" ; echo built $tag; }
build mb3 -DCB_UNITS_MIN_BLOCKS=3 &
build mb4 -DCB_UNITS_MIN_BLOCKS=4 &
build u2_mb4 -DCB_UNITS_MIN_BLOCKS=4 -DCB_LW_UMAX=2 -DCB_SW_UMAX=2 &
build u2_mb6 -DCB_UNITS_MIN_BLOCKS=6 -DCB_LW_UMAX=2 -DCB_SW_UMAX=2 &
build u2_mb8 -DCB_UNITS_MIN_BLOCKS=8 -DCB_LW_UMAX=2 -DCB_SW_UMAX=2 &
wait
ls -la ../variants

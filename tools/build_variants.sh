#!/bin/bash
# Build flag variants of the engine for on-GPU A/B timing (development tool; outputs under variants/, git-ignored *.so)
cd "$(dirname "$0")/../climt_b200" || exit 1
mkdir -p ../variants
rm -f ../variants/*.so
BASE="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --expt-relaxed-constexpr -Xcompiler -fPIC -shared --fmad=true"
build() { tag=$1; shift; nvcc $BASE "$@" -o ../variants/libclimt_b200_$tag.so csrc/lw_engine.cu csrc/sw_engine.cu csrc/gray_engine.cu csrc/cork_engine.cu -lcudart -ldl 2>&1 | grep -E "error" ; echo built $tag; }
build sw5_lw8 -DCB_LW_RT_MIN_BLOCKS=8 -DCB_SW_RT_MIN_BLOCKS=5 &
build sw6_lw5 -DCB_LW_RT_MIN_BLOCKS=5 -DCB_SW_RT_MIN_BLOCKS=6 &
build sw3_tau5 -DCB_SW_RT_MIN_BLOCKS=3 -DCB_LW_TAU_MIN_BLOCKS=5 -DCB_SW_TAU_MIN_BLOCKS=5 &
build z8 -DCB_LW_LAYER_CHUNKS=8 -DCB_SW_LAYER_CHUNKS=8 &
wait
ls -la ../variants

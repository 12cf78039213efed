#!/bin/bash
# Build flag variants of the engine for on-GPU A/B timing (development tool; outputs under variants/, git-ignored *.so)
cd "$(dirname "$0")/../climt_b200" || exit 1
mkdir -p ../variants
rm -f ../variants/*.so
BASE="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --expt-relaxed-constexpr -Xcompiler -fPIC -shared --fmad=true"
build() { tag=$1; shift; nvcc $BASE "$@" -o ../variants/libclimt_b200_$tag.so csrc/lw_engine.cu csrc/sw_engine.cu csrc/gray_engine.cu csrc/cork_engine.cu csrc/marshal.cu csrc/emanuel_engine.cu csrc/adjacent_engine.cu csrc/simple_physics.cu -lcudart -ldl 2>&1 | grep -E "error" ; echo built $tag; }
# candidates not yet measured (r01 left off at: SW transfer 1 g-point / 8 blocks = 64 registers with 164 B of spills)
build sw_b7 -DCB_SW_RT_MIN_BLOCKS=7 &
build sw_b9 -DCB_SW_RT_MIN_BLOCKS=9 &
build lw_b7 -DCB_LW_RT_MIN_BLOCKS=7 &
build tau8 -DCB_LW_LAYER_CHUNKS=8 -DCB_SW_LAYER_CHUNKS=8 &
wait
ls -la ../variants

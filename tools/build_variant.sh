#!/bin/bash
# Development helper: build libclimt_b200_<tag>.so with extra -D flags for A/B measurements on the GPU box
#   tools/build_variant.sh <tag> [-DCB_LW_STAGE=0 ...]      then   CLIMT_B200_SO=$PWD/climt_b200/libclimt_b200_<tag>.so python bench.py
# Objects are compiled in parallel and cached per flag set under build/.
set -e
tag=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
src=$root/climt_b200/csrc
key=$(echo "$@" | md5sum | cut -c1-8)
obj=$root/build/obj_$key
mkdir -p "$obj"
flags="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --expt-relaxed-constexpr -Xcompiler -fPIC --fmad=true"
pids=()
for f in lw_engine sw_engine gray_engine cork_engine marshal emanuel_engine adjacent_engine simple_physics; do
  if [ ! -f "$obj/$f.o" ] || [ -n "$(find "$src" "$root/include" -newer "$obj/$f.o" -type f | head -1)" ]; then
    nvcc $flags "$@" -c "$src/$f.cu" -o "$obj/$f.o" &
    pids+=($!)
  fi
done
for p in "${pids[@]}"; do wait "$p"; done
nvcc -shared -o "$root/climt_b200/libclimt_b200_$tag.so" "$obj"/*.o -lcudart -ldl
echo "$root/climt_b200/libclimt_b200_$tag.so"

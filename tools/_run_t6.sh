for v in t512 t1024; do
CLIMT_B200_SO=$PWD/climt_b200/libclimt_b200_$v.so SKIP_TESTS=1 TILES=1 KERNELS=none bash tools/tile_check.sh r2t6_$v 2>&1 | grep -E "^tile|^bench"
done

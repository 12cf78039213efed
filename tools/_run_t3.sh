SKIP_TESTS=1 TILES=1 bash tools/tile_check.sh r2t3
CLIMT_B200_SO=$PWD/climt_b200/libclimt_b200_t512.so SKIP_TESTS=1 TILES=1 bash tools/tile_check.sh r2t3_512

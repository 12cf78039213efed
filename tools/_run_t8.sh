TESTS="tests/test_lw_gpu.py tests/test_northstar_shape_gpu.py tests/test_mcica_symbols_gpu.py" TILES=1 bash tools/tile_check.sh r2t8
bash tools/_run_t7.sh

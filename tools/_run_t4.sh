O=gpurun_out/r2t4; mkdir -p $O
export CLIMT_B200_SO=$PWD/climt_b200/libclimt_b200_t256.so
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_lw_tile -c 1 -o /tmp/tile -f python tools/time_engine.py > $O/ncu.log 2>&1
ncu -i /tmp/tile.ncu-rep --page details > $O/details.txt 2>&1
ncu -i /tmp/tile.ncu-rep --page raw --csv > $O/raw.csv 2>&1
ncu -i /tmp/tile.ncu-rep --page source --csv > $O/source.csv 2>&1
python tools/ncu_top_source.py $O/source.csv 40 > $O/top_source.csv 2>&1
ls -la $O

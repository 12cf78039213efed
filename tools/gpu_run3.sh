set -x
O=gpurun_out/r3
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log
tail -8 $O/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke exit $?" >> $O/smoke.log; tail -5 $O/smoke.log
timeout 600 python bench.py > $O/bench.log 2>&1; tail -2 $O/bench.log
for u in 4 8 2; do CLIMT_B200_CORK_U=$u timeout 600 python tools/time_cork.py 2>&1 | tail -1 >> $O/cork.jsonl; done
BANDS=0 timeout 600 python tools/time_cork.py 2>&1 | tail -1 >> $O/cork.jsonl
NCOL=8192 timeout 600 python tools/time_cork.py 2>&1 | tail -1 >> $O/cork.jsonl
cat $O/cork.jsonl
for k in k_sw_units k_units; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -o /tmp/$k -f python tools/time_engine.py > $O/ncu_$k.log 2>&1
  ncu -i /tmp/$k.ncu-rep --page details > $O/${k}_details.txt 2>&1
  ncu -i /tmp/$k.ncu-rep --page raw --csv > $O/${k}_raw.csv 2>&1
  ncu -i /tmp/$k.ncu-rep --page source --csv > /tmp/${k}_source.csv 2>&1
  python tools/ncu_top_source.py /tmp/${k}_source.csv 80 > $O/${k}_source_top.csv 2>$O/${k}_source_top.err
done
NCOL=16384 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_cork_units -c 1 -o /tmp/cork -f python tools/time_cork.py > $O/ncu_cork.log 2>&1
ncu -i /tmp/cork.ncu-rep --page details > $O/k_cork_units_details.txt 2>&1
ncu -i /tmp/cork.ncu-rep --page raw --csv > $O/k_cork_units_raw.csv 2>&1
ncu -i /tmp/cork.ncu-rep --page source --csv > /tmp/cork_source.csv 2>&1
python tools/ncu_top_source.py /tmp/cork_source.csv 60 > $O/k_cork_units_source_top.csv 2>$O/cork_source_top.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 1 > $O/ncu_bench.log 2>&1
ls -la $O; du -sh gpurun_out

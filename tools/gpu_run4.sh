set -x
O=gpurun_out/r4
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log
tail -15 $O/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke exit $?" >> $O/smoke.log; tail -6 $O/smoke.log
timeout 600 python bench.py > $O/bench.log 2>&1; tail -2 $O/bench.log
for v in default rt4 rt5 rt8_6 u4 z1 z2_tau3 tauu2; do
  if [ $v = default ]; then unset CLIMT_B200_SO; else export CLIMT_B200_SO=$PWD/variants/libclimt_b200_$v.so; fi
  timeout 300 python tools/time_engine.py 2>&1 | tail -1 >> $O/variants.jsonl
done
unset CLIMT_B200_SO
CLOUDS=1 timeout 300 python tools/time_engine.py 2>&1 | tail -1 >> $O/variants.jsonl
cat $O/variants.jsonl
for u in 4 8 2; do CLIMT_B200_CORK_U=$u timeout 600 python tools/time_cork.py 2>&1 | tail -1 >> $O/cork.jsonl; done
BANDS=0 timeout 600 python tools/time_cork.py 2>&1 | tail -1 >> $O/cork.jsonl
NCOL=8192 timeout 600 python tools/time_cork.py 2>&1 | tail -1 >> $O/cork.jsonl
cat $O/cork.jsonl
for k in k_sw_transfer k_sw_taumol k_units k_lw_taumol; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -o /tmp/$k -f python tools/time_engine.py > $O/ncu_$k.log 2>&1
  ncu -i /tmp/$k.ncu-rep --page details > $O/${k}_details.txt 2>&1
  ncu -i /tmp/$k.ncu-rep --page raw --csv > $O/${k}_raw.csv 2>&1
done
NCOL=16384 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_cork_units -c 1 -o /tmp/cork -f python tools/time_cork.py > $O/ncu_cork.log 2>&1
ncu -i /tmp/cork.ncu-rep --page details > $O/k_cork_units_details.txt 2>&1
ncu -i /tmp/cork.ncu-rep --page raw --csv > $O/k_cork_units_raw.csv 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 1 > $O/ncu_bench.log 2>&1
ls -la $O; du -sh gpurun_out

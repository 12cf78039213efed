set -x
O=gpurun_out/r2
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log
tail -5 $O/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke exit $?" >> $O/smoke.log; tail -2 $O/smoke.log
timeout 600 python bench.py > $O/bench.log 2>&1; tail -2 $O/bench.log
for v in default mb3 mb4 u2_mb4 u2_mb6 u2_mb8; do
  if [ $v = default ]; then unset CLIMT_B200_SO; else export CLIMT_B200_SO=$PWD/variants/libclimt_b200_$v.so; fi
  timeout 300 python tools/time_engine.py 2>&1 | tail -1 >> $O/variants.jsonl
done
unset CLIMT_B200_SO
CLOUDS=1 timeout 300 python tools/time_engine.py 2>&1 | tail -1 >> $O/variants.jsonl
cat $O/variants.jsonl
for k in k_sw_units k_units; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:^$k -c 1 -o /tmp/$k -f python tools/time_engine.py > $O/ncu_$k.log 2>&1
  ncu -i /tmp/$k.ncu-rep --page details > $O/${k}_details.txt 2>&1
  ncu -i /tmp/$k.ncu-rep --page raw --csv > $O/${k}_raw.csv 2>&1
  ncu -i /tmp/$k.ncu-rep --page source --csv > $O/${k}_source.csv 2>&1
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 1 > $O/ncu_bench.log 2>&1
ls -la $O; du -sh gpurun_out

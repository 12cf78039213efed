set -x
O=gpurun_out/r8
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log; tail -3 $O/pytest.log
timeout 600 python bench.py > $O/bench.log 2>&1; tail -1 $O/bench.log
timeout 300 python tools/time_engine.py 2>&1 | tail -1 >> $O/variants.jsonl
CLOUDS=1 timeout 300 python tools/time_engine.py 2>&1 | tail -1 >> $O/variants.jsonl
MCICA=1 NLAY=72 NCOL=16384 timeout 300 python tools/time_engine.py 2>&1 | tail -1 >> $O/variants.jsonl
cat $O/variants.jsonl
MCICA=1 NLAY=72 NCOL=16384 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_mcica.csv python tools/time_engine.py > $O/ncu_mcica.log 2>&1
CLOUDS=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_cloudy.csv python tools/time_engine.py > $O/ncu_cloudy.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_bench.log 2>&1
for k in k_sw_transfer k_units k_lw_taumol; do
  timeout 900 ncu --set full --clock-control none -k regex:$k -c 1 -o /tmp/$k -f python tools/time_engine.py > $O/ncu_$k.log 2>&1
  ncu -i /tmp/$k.ncu-rep --page details > $O/${k}_details.txt 2>&1
  ncu -i /tmp/$k.ncu-rep --page raw --csv > $O/${k}_raw.csv 2>&1
done

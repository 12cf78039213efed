set -x
O=gpurun_out/r10
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log; tail -3 $O/pytest.log
timeout 600 python bench.py --no-cpu-baseline > $O/bench.log 2>&1; tail -1 $O/bench.log
timeout 300 python tools/time_engine.py 2>&1 | tail -1 >> $O/variants.jsonl
CLOUDS=1 timeout 300 python tools/time_engine.py 2>&1 | tail -1 >> $O/variants.jsonl
MCICA=1 NLAY=72 NCOL=16384 timeout 300 python tools/time_engine.py 2>&1 | tail -1 >> $O/variants.jsonl
cat $O/variants.jsonl
timeout 900 ncu --set full --clock-control none -k regex:k_units -c 1 -o /tmp/k_units -f python tools/time_engine.py > $O/ncu_k_units.log 2>&1
ncu -i /tmp/k_units.ncu-rep --page details > $O/k_units_details.txt 2>&1
ncu -i /tmp/k_units.ncu-rep --page raw --csv > $O/k_units_raw.csv 2>&1

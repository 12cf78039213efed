set -x
O=gpurun_out/r6
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log; tail -4 $O/pytest.log
timeout 600 python bench.py > $O/bench.log 2>&1; tail -1 $O/bench.log
timeout 600 python tools/e2e_probe.py > $O/e2e_probe.log 2>&1; tail -1 $O/e2e_probe.log
timeout 300 python tools/time_engine.py 2>&1 | tail -1 >> $O/variants.jsonl
CLOUDS=1 timeout 300 python tools/time_engine.py 2>&1 | tail -1 >> $O/variants.jsonl
cat $O/variants.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_bench.log 2>&1

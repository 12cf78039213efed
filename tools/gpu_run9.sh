set -x
O=gpurun_out/r9
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log; tail -3 $O/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke exit $?" >> $O/smoke.log; tail -5 $O/smoke.log
timeout 600 python bench.py > $O/bench.log 2>&1; tail -1 $O/bench.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.log 2>&1; tail -1 $O/bench_ref.log
timeout 300 python tools/time_engine.py 2>&1 | tail -1 >> $O/variants.jsonl
CLOUDS=1 timeout 300 python tools/time_engine.py 2>&1 | tail -1 >> $O/variants.jsonl
MCICA=1 NLAY=72 NCOL=16384 timeout 300 python tools/time_engine.py 2>&1 | tail -1 >> $O/variants.jsonl
cat $O/variants.jsonl
timeout 600 python tools/time_cork.py 2>&1 | tail -1 >> $O/cork.jsonl
cat $O/cork.jsonl
nproc; lscpu | grep "Model name"

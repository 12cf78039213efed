set -x
O=gpurun_out/r7
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log; tail -4 $O/pytest.log
timeout 600 python bench.py > $O/bench.log 2>&1; tail -1 $O/bench.log
for v in default sw5_lw8 sw6_lw5 sw3_tau5 z8; do
  if [ $v = default ]; then unset CLIMT_B200_SO; else export CLIMT_B200_SO=$PWD/variants/libclimt_b200_$v.so; fi
  timeout 300 python tools/time_engine.py 2>&1 | tail -1 >> $O/variants.jsonl
done
unset CLIMT_B200_SO
CLOUDS=1 timeout 300 python tools/time_engine.py 2>&1 | tail -1 >> $O/variants.jsonl
MCICA=1 NLAY=72 NCOL=16384 timeout 300 python tools/time_engine.py 2>&1 | tail -1 >> $O/variants.jsonl
NCOL=65536 timeout 300 python tools/time_engine.py 2>&1 | tail -1 >> $O/variants.jsonl
cat $O/variants.jsonl
timeout 600 python tools/e2e_probe.py > $O/e2e_probe.log 2>&1; tail -1 $O/e2e_probe.log
for u in 4 8; do CLIMT_B200_CORK_U=$u timeout 600 python tools/time_cork.py 2>&1 | tail -1 >> $O/cork.jsonl; done
cat $O/cork.jsonl

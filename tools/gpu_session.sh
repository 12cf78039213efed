#!/bin/bash
# One GPU-box session of the development loop (run as: gpurun --timeout 2400 -- 'bash tools/gpu_session.sh <tag>').
# Writes small text artefacts only under gpurun_out/<tag>/ (the .ncu-rep files stay in /tmp on the box: gpurun_out is capped at
# 64 MiB); copy what should be judged into profiles/.
#   parity tests -> smoke -> bench line -> per-kernel launch list of the bench command -> one `ncu --set full` capture per hot
#   kernel, exported as details + raw CSV.
set -x
O=gpurun_out/${1:-session}
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log; tail -3 $O/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke exit $?" >> $O/smoke.log; tail -5 $O/smoke.log
timeout 600 python bench.py > $O/bench.log 2>&1; tail -1 $O/bench.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.log 2>&1; tail -1 $O/bench_ref.log
timeout 600 python bench_extra.py --workload mcica > $O/bench_mcica.log 2>&1; tail -1 $O/bench_mcica.log
timeout 600 python bench_extra.py --workload cork > $O/bench_cork.log 2>&1; tail -1 $O/bench_cork.log
timeout 600 python bench_extra.py --workload gmd > $O/bench_gmd.log 2>&1; tail -1 $O/bench_gmd.log
timeout 200 python tools/pipe_trace.py > $O/pipe_trace.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_bench.log 2>&1
for k in k_sw_transfer k_sw_taumol k_lw_tile k_lw_taumol; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -o /tmp/$k -f python tools/time_engine.py > $O/ncu_$k.log 2>&1
  ncu -i /tmp/$k.ncu-rep --page details > $O/${k}_details.txt 2>&1
  ncu -i /tmp/$k.ncu-rep --page raw --csv > $O/${k}_raw.csv 2>&1
done
NCOL=16384 timeout 900 ncu --set full --clock-control none -k regex:k_cork_units -c 1 -o /tmp/cork -f python tools/time_cork.py > $O/ncu_cork.log 2>&1
ncu -i /tmp/cork.ncu-rep --page details > $O/k_cork_units_details.txt 2>&1
ncu -i /tmp/cork.ncu-rep --page raw --csv > $O/k_cork_units_raw.csv 2>&1
ls -la $O; du -sh gpurun_out

#!/bin/bash
# Short GPU-box session: parity tests -> smoke -> bench line -> launch list of the bench command -> `ncu --set full` capture of the
# four hot RRTMG kernels (details + raw CSV).   gpurun --timeout 2000 -- 'bash tools/gpu_session_short.sh <tag>'
set -x
O=gpurun_out/${1:-session}
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log; tail -3 $O/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke exit $?" >> $O/smoke.log; tail -5 $O/smoke.log
timeout 600 python bench.py > $O/bench.log 2>&1; tail -1 $O/bench.log
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.log 2>&1; tail -1 $O/bench_ref.log
timeout 100 python tools/time_engine.py >> $O/time.jsonl 2>>$O/err.log
MCICA=1 NCOL=16384 NLAY=72 timeout 100 python tools/time_engine.py >> $O/time.jsonl 2>>$O/err.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > $O/ncu_bench.log 2>&1
for k in ${KERNELS:-k_sw_transfer k_sw_taumol k_lw_tile k_lw_taumol}; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -o /tmp/$k -f python tools/time_engine.py > $O/ncu_$k.log 2>&1
  ncu -i /tmp/$k.ncu-rep --page details > $O/${k}_details.txt 2>&1
  ncu -i /tmp/$k.ncu-rep --page raw --csv > $O/${k}_raw.csv 2>&1
done
ls -la $O; du -sh gpurun_out

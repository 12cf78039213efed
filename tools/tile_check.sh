#!/bin/bash
# Development helper (GPU box): the column-tile transfer kernels -- canary under a short timeout (named-barrier hand-over: a
# protocol error hangs), the radiation GPU tests, timings against the unit form, DRAM bytes of one launch.
#   gpurun --timeout 900 -- 'bash tools/tile_check.sh <outdir>'
out=gpurun_out/${1:-tile}
mkdir -p $out
timeout 120 python tools/time_engine.py > $out/canary.json 2>$out/canary.err; echo "canary rc=$?" | tee -a $out/canary.err
grep -q "canary rc=0" $out/canary.err || { tail -5 $out/canary.err; exit 1; }
if [ -z "$SKIP_TESTS" ]; then
(timeout 600 python -m pytest ${TESTS:-tests/test_lw_gpu.py tests/test_sw_gpu.py tests/test_mcica_symbols_gpu.py tests/test_northstar_shape_gpu.py tests/test_device_state_gpu.py tests/test_host_pipeline_gpu.py tests/test_humid_crosscheck.py} -m gpu -q 2>&1 | tail -12) > $out/pytest.log
cat $out/pytest.log
fi
for t in ${TILES:-1 0}; do
  for cfg in "0 0 8192 60" "0 1 8192 60" "1 0 16384 72"; do
    set -- $cfg
    CLIMT_B200_LW_TILE=$t CLIMT_B200_SW_SCAN=$t MCICA=$1 CLOUDS=$2 NCOL=$3 NLAY=$4 timeout 120 python tools/time_engine.py 2>>$out/err.log | sed "s/^{/{\"tile\": $t, /" >> $out/time.jsonl
  done
done
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active
for k in ${KERNELS:-k_lw_tile}; do
  timeout 300 ncu --metrics $M --clock-control none -k regex:$k -c 1 --csv --log-file $out/ncu_$k.csv python tools/time_engine.py > /dev/null 2>>$out/err.log
  MCICA=1 NCOL=16384 NLAY=72 timeout 300 ncu --metrics $M --clock-control none -k regex:$k -c 1 --csv --log-file $out/ncu_mcica_$k.csv python tools/time_engine.py > /dev/null 2>>$out/err.log
done
timeout 200 python bench.py --steps 10 --no-cpu-baseline --no-extras > $out/bench.json 2>>$out/err.log
python - <<PY
import json, csv, glob
for l in open("$out/time.jsonl"):
    d = json.loads(l)
    print("tile", d["tile"], "mcica" if d["mcica"] else ("cloudy" if d["clouds"] else "clear"), "sw_units %.3f lw_units %.3f | sw_step %.3f lw_step %.3f" % (d["sw_units_ms"], d["lw_units_ms"], d["sw_step_ms"], d["lw_step_ms"]), d["lw_checksum"])
for f in sorted(glob.glob("$out/ncu_*.csv")):
    rows = [r for r in csv.reader(open(f)) if len(r) > 10]
    print(f.split("/")[-1], " | ".join("%s=%s" % (r[-3].split("__")[1][:22], r[-1]) for r in rows[1:]))
try:
    d = json.load(open("$out/bench.json")); print("bench", d["value"], d["ms_per_step"], d["e2e"]["value"])
except Exception as e: print("bench failed", e)
PY
tail -3 $out/err.log

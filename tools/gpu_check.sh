#!/bin/bash
# Development helper (GPU box): canary for the ragged-chunk barrier, then the radiation GPU tests, kernel timings and a short bench.
#   gpurun --timeout 600 -- 'bash tools/gpu_check.sh <outdir>'
out=gpurun_out/${1:-check}
mkdir -p $out
timeout 90 python tools/canary.py > $out/canary.log 2>&1; echo "canary rc=$?" >> $out/canary.log
tail -12 $out/canary.log
grep -q "canary rc=0" $out/canary.log || exit 1
(timeout 300 python -m pytest tests/test_lw_gpu.py tests/test_sw_gpu.py tests/test_mcica_symbols_gpu.py tests/test_northstar_shape_gpu.py tests/test_device_state_gpu.py tests/test_host_pipeline_gpu.py -m gpu -x -q 2>&1 | tail -8) > $out/pytest.log
timeout 100 python tools/time_engine.py >> $out/time.jsonl 2>>$out/err.log
MCICA=1 NCOL=16384 NLAY=72 timeout 100 python tools/time_engine.py >> $out/time.jsonl 2>>$out/err.log
timeout 200 python bench.py --steps 10 --no-cpu-baseline --no-extras > $out/bench.json 2>>$out/err.log
cat $out/pytest.log
python - <<PY
import json
for l in open("$out/time.jsonl"):
    d = json.loads(l)
    print(d["mcica"], "sw_step %.3f units %.3f taumol %.3f | lw_step %.3f units %.3f taumol %.3f" % (d["sw_step_ms"], d["sw_units_ms"], d["sw_taumol_ms"], d["lw_step_ms"], d["lw_units_ms"], d["lw_taumol_ms"]), d["lw_checksum"])
d = json.load(open("$out/bench.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"])
PY
tail -3 $out/err.log

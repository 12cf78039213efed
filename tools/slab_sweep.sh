#!/bin/bash
# Development helper (GPU box): time the slab form of the two transfer kernels against the default form for several
# blocks-per-SM settings, clear sky 8192 x 60 and McICA 16384 x 72, and read the DRAM bytes of one launch each under ncu.
#   gpurun --timeout 900 -- 'bash tools/slab_sweep.sh <outdir> "<so tag>:<lw bps>/<sw bps> ..." ...'
#   e.g.  "default:0/0 2/1 3/2 4/3" "lwu1:1/0 2/0"
out=gpurun_out/${1:-slab}; shift
mkdir -p $out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active
for spec in "$@"; do
  tag=${spec%%:*}; list=${spec#*:}
  so=$PWD/climt_b200/libclimt_b200.so
  [ "$tag" != default ] && so=$PWD/climt_b200/libclimt_b200_$tag.so
  for pair in $list; do
    lwb=${pair%%/*}; swb=${pair#*/}
    export CLIMT_B200_SO=$so CLIMT_B200_LW_SLAB=$lwb CLIMT_B200_SW_SLAB=$swb
    for cfg in "0 8192 60" "1 16384 72"; do
      set -- $cfg
      MCICA=$1 NCOL=$2 NLAY=$3 timeout 120 python tools/time_engine.py 2>>$out/err.log | sed "s/^{/{\"lw_bps\": $lwb, \"sw_bps\": $swb, \"build\": \"$tag\", /" >> $out/time.jsonl
    done
    [ -n "$NO_NCU" ] && continue
    for k in k_units k_sw_transfer; do
      timeout 300 ncu --metrics $M --clock-control none -k regex:$k -c 1 --csv --log-file $out/ncu_${tag}_${lwb}_${swb}_$k.csv python tools/time_engine.py > /dev/null 2>>$out/err.log
    done
  done
done
python - <<PY
import json, csv, glob
for l in open("$out/time.jsonl"):
    d = json.loads(l)
    print(d["build"], "lw", d["lw_bps"], "sw", d["sw_bps"], "mcica" if d["mcica"] else "clear", "sw_units %.3f lw_units %.3f | sw_step %.3f lw_step %.3f" % (d["sw_units_ms"], d["lw_units_ms"], d["sw_step_ms"], d["lw_step_ms"]), d["lw_checksum"])
for f in sorted(glob.glob("$out/ncu_*.csv")):
    rows = [r for r in csv.reader(open(f)) if len(r) > 10]
    print(f.split("/")[-1], " | ".join("%s=%s" % (r[-3].split("__")[1][:22], r[-1]) for r in rows[1:]))
PY
tail -3 $out/err.log

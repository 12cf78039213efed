for v in skipc skipp s1024 s768; do
  for cfg in "0 0 8192 60" "0 1 8192 60"; do
    set -- $cfg
    CLIMT_B200_SO=$PWD/climt_b200/libclimt_b200_$v.so MCICA=$1 CLOUDS=$2 NCOL=$3 NLAY=$4 timeout 120 python tools/time_engine.py 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v', 'mcica' if d['mcica'] else ('cloudy' if d['clouds'] else 'clear'), 'sw_units %.3f lw_units %.3f' % (d['sw_units_ms'], d['lw_units_ms']))"
  done
done

#!/bin/bash
# usage: variants.sh name1 name2 ...   (libisy_<name>.so in invertsy_b200/)
for v in "$@"; do
  ISY_B200_LIB=$PWD/invertsy_b200/libisy_$v.so timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/var_$v.json 2> gpurun_out/var_$v.err || echo "FAILED $v"
  python - "$v" <<'PY'
import json,sys
v=sys.argv[1]
try:
    d=json.loads(open(f'gpurun_out/var_{v}.json').read().strip().splitlines()[-1])
    print(v, round(d['ms_per_step'],3), d.get('phases'), d['clocks']['sm_mhz'], d['check'])
except Exception as e: print(v,'ERR',e)
PY
done

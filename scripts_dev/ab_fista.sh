#!/bin/bash
for v in "$@"; do
  echo "== $v"
  env $v python bench.py --algo fista --configs none --no-cpu-baseline --steps 3 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('recon/s %.2f  ms/iter %.4f  phases' % (d['value'], d['ms_per_step']/300), {k: round(v*1e3,1) for k,v in d['roofline']['phases_ms_per_iter'].items()}, 'clk', d['clocks']['sm_mhz'])
"
done

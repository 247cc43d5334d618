#!/bin/bash
B="timeout 90 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
run() { echo "== $*"; env BDS_TRK_TIMING=1 $B "$@" 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('x_rt',round(d['config']['x_realtime'],1),'kernel_ms',round(d['roofline']['kernel_ms_per_launch'],2))
    elif 'closure' in l: print(l.rstrip()[:300])
"; }
run --channels 60
run --channels 30
run --channels 15
run --channels 8

#!/bin/bash
# developer sweep of the tracking kernel's tuning knobs on one B200 (run under gpurun); short timeouts on purpose
B="timeout 90 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
run() { echo "== $*"; env "$@" $B 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('x_rt',round(d['config']['x_realtime'],1),'kernel_ms',round(d['roofline']['kernel_ms_per_launch'],2))
    elif 'bds timing' in l: print(l.rstrip()[:300])
"; }
run BDS_TRK_TIMING=1
for pa in 1 2 3; do run BDS_TRK_PASSES=$pa; done
for ah in 0 1 2 3; do run BDS_TRK_AHEAD=$ah; done     # clamped to stages - 1 by the library

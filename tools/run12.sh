#!/bin/bash
timeout 200 python -m pytest tests/test_gpu_tracking.py -m gpu -x -q 2>&1 | tail -3
B="timeout 90 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
run() { echo "== $*"; env "$@" $B 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('x_rt',round(d['config']['x_realtime'],1),'kernel_ms',round(d['roofline']['kernel_ms_per_launch'],2))
"; }
run A=1
run BDS_TRK_PASSES=3
run BDS_TRK_AHEAD=3
run BDS_TRK_AHEAD=1
timeout 90 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --channels 8 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('8ch x_rt',round(d['config']['x_realtime'],1))
"

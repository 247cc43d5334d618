#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tracking.py -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest.log
B="timeout 90 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
run() { echo "== $*"; env "$@" $B 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('x_rt',round(d['config']['x_realtime'],1),'kernel_ms',round(d['roofline']['kernel_ms_per_launch'],2))
    else: print(l.rstrip()[:400])
"; }
run BDS_TRK_TIMING=1
run BDS_TRK_TUNE=8
run BDS_TRK_PASSES=3
run BDS_TRK_AHEAD=1
run BDS_TRK_AHEAD=3
run BDS_TRK_STAGES=3
